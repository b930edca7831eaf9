"""Does interleaving independent sub-batches (separate engines / streams) beat one big batch?"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bvhar_b200._engine import CudaFit
from bvhar_b200._partition import shard_fit_kwargs, shard_units
from bvhar_b200._abi import PackedFit
total = 1024
K = 20
m = bench.build_model(total, 64, 0, seed=1, device=0)
kw = dict(num_chains=total, num_iter=64, num_burn=0, thin=1, x=m.design_, y=m.response_, param_cov=m.cov_spec_.to_dict(),
          param_prior=m.spec_.to_dict(), param_intercept=m.intercept_spec_.to_dict(), param_init=m.init_, prior_type=int(m._prior_type),
          grp_id=m._group_id, own_id=m._own_id, cross_id=m._cross_id, grp_mat=m.group_, include_mean=True, seed_chain=m._seeds(1, total),
          is_group=True, record_mask=0)
for parts in (1, 2, 3, 4):
    fits = [CudaFit(PackedFit(**shard_fit_kwargs(kw, shard_units(1, total, parts, r)))) for r in range(parts)]
    for f in fits: f.step(3, False)
    t0 = time.perf_counter()
    for f in fits: f.step(K, False, sync=False)
    for f in fits: f.sync()
    dt = time.perf_counter() - t0
    print(f"{parts} engine(s): {1e3 * dt / K:.3f} ms per sweep of all {total} chains -> {total * K / dt:.0f} chain-sweeps/s", flush=True)
    for f in fits: f.close()
