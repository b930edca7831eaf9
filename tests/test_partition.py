"""Multi-GPU host logic on CPU: sharding of (window, chain) units, invariance of the draws to the number of
shards (Philox keyed on the global unit id), and the final gather over a world_size-2 gloo group.
The oracle stands in for the per-rank engine (both implement the same addressed stream; the CUDA engine
is checked against it bit-for-bit in tests/test_gpu_parity.py and re-checked for shard invariance on the
GPU in tests/test_gpu_scale.py)."""
import os
import socket
import sys

import numpy as np
import pytest
from numpy.testing import assert_array_equal

import oracle_binding as ob
import problems as pr
from bvhar_b200._abi import PackedFit
from bvhar_b200._partition import Shard, gather_units, merge_shards, shard_fit_kwargs, shard_units


@pytest.mark.parametrize("W,C,world", [(1, 8, 2), (1, 5, 4), (1, 3, 8), (7, 4, 2), (10, 2, 8), (3, 2, 4)])
def test_shards_tile_the_unit_range(W, C, world):
    seen = []
    for r in range(world):
        s = shard_units(W, C, world, r)
        assert s.by == ("chain" if W == 1 else "window")
        seen += list(range(s.unit_id_base, s.unit_id_base + s.n_units))
    assert seen == list(range(W * C))
    sizes = [shard_units(W, C, world, r).n_units for r in range(world)]
    assert max(sizes) - min(sizes) <= (1 if W == 1 else C)


def test_shard_argument_checks():
    with pytest.raises(ValueError):
        shard_units(1, 4, 2, 2)
    with pytest.raises(ValueError):
        shard_units(0, 4, 1, 0)


def _run_sharded(kw, W, C, world):
    parts = {"alpha_record": [], "a_record": []}
    for r in range(world):
        sh = shard_units(W, C, world, r)
        skw = shard_fit_kwargs(kw, sh)
        if skw is None:
            continue
        f = ob.OracleFit(PackedFit(**skw)); f.run()
        for nm in parts:
            parts[nm].append(np.stack([f.record(nm, w, c) for w in range(len(sh.windows)) for c in range(len(sh.chains))]))
        f.close()
    return {nm: merge_shards(v) for nm, v in parts.items()}


@pytest.mark.parametrize("world", [2, 3])
def test_chain_sharding_reproduces_the_unsharded_draws(world):
    kw, _ = pr.make_problem(prior="Horseshoe", sv=True, dim=2, T=30, chains=5, iters=4, burn=1)
    one = _run_sharded(kw, 1, 5, 1)
    many = _run_sharded(kw, 1, 5, world)
    for nm in one:
        assert_array_equal(one[nm], many[nm])


def test_window_sharding_reproduces_the_unsharded_draws():
    wins = [(0, 24), (1, 24), (2, 24), (3, 24), (4, 24)]
    kw, _ = pr.make_problem(prior="SSVS", sv=False, dim=2, T=32, chains=2, iters=3, burn=0, windows=wins)
    one = _run_sharded(kw, 5, 2, 1)
    many = _run_sharded(kw, 5, 2, 2)
    for nm in one:
        assert_array_equal(one[nm], many[nm])


# ---------------------------------------------------------------- world_size-2 gloo group
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, kw, W, C, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = shard_units(W, C, world, rank)
        skw = shard_fit_kwargs(kw, sh)
        local = None
        if skw is not None:
            f = ob.OracleFit(PackedFit(**skw)); f.run()
            local = np.stack([f.record("alpha_record", w, c) for w in range(len(sh.windows)) for c in range(len(sh.chains))])
        got = gather_units(local, sh)
        if rank == 0:
            q.put(got)
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("W,C", [(1, 3), (3, 2)])
def test_gloo_world2_gather_matches_single_process(W, C):
    import torch.multiprocessing as mp
    wins = None if W == 1 else [(i, 24) for i in range(W)]
    kw, _ = pr.make_problem(prior="DL", sv=True, dim=2, T=32, chains=C, iters=3, burn=0, windows=wins)
    ref = _run_sharded(kw, W, C, 1)["alpha_record"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kw, W, C, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert_array_equal(got, ref)
