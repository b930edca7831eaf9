"""Sharding of (window, chain) sampling units over the GPUs of one box.

Chains never interact (triangular.h:1222-1226) and rolling / expanding windows never interact
(forecaster.h:619-634), so the unit of distribution is the (window, chain) pair and there is NO
collective on the sampling path: every rank runs its own engine on its own GPU, and the only
exchange is the final gather of records / forecasts.

Shards are contiguous blocks of the global unit order ``u = window * n_chains + chain``:

* one window (a plain ``McmcRun`` fit)  -> chains are split, ``unit_id_base`` = first chain;
* several windows (roll / expand)       -> whole windows are split (the chains of a window share one
  device copy of the design), ``unit_id_base`` = first window * n_chains.

Because every draw is addressed by the Philox key ``(seed_chain[u], unit_id_base + u_local)``, a
unit's trajectory is bit-identical for any number of shards.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    by: str          # "chain" or "window"
    start: int       # first chain (by == "chain") or first window (by == "window") of the shard
    count: int       # number of chains / windows of the shard (may be 0)
    n_windows: int   # windows of the whole job
    n_chains: int    # chains per window of the whole job

    @property
    def unit_id_base(self) -> int:
        return self.start if self.by == "chain" else self.start * self.n_chains

    @property
    def n_units(self) -> int:
        return self.count if self.by == "chain" else self.count * self.n_chains

    @property
    def windows(self) -> range:
        return range(self.n_windows) if self.by == "chain" else range(self.start, self.start + self.count)

    @property
    def chains(self) -> range:
        return range(self.start, self.start + self.count) if self.by == "chain" else range(self.n_chains)


def _block(total: int, world: int, rank: int):
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def shard_units(n_windows: int, n_chains: int, world: int, rank: int) -> Shard:
    """Contiguous block partition; earlier ranks take the remainder."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("need 0 <= rank < world")
    if n_windows < 1 or n_chains < 1:
        raise ValueError("empty job")
    if n_windows == 1:
        s, c = _block(n_chains, world, rank)
        return Shard(rank, world, "chain", s, c, n_windows, n_chains)
    s, c = _block(n_windows, world, rank)
    return Shard(rank, world, "window", s, c, n_windows, n_chains)


def shard_fit_kwargs(kw: Dict, shard: Shard) -> Optional[Dict]:
    """Restrict the keyword arguments of ``PackedFit`` (= ``McmcRun``'s arguments + the window table)
    to one shard.  Returns None for an empty shard."""
    if shard.n_units == 0:
        return None
    out = dict(kw)
    seeds = np.asarray(kw["seed_chain"]).reshape(shard.n_windows, shard.n_chains)
    if shard.by == "chain":
        sl = slice(shard.start, shard.start + shard.count)
        out["num_chains"] = shard.count
        out["param_init"] = list(kw["param_init"])[sl]
        out["seed_chain"] = seeds[:, sl]
    else:
        sl = slice(shard.start, shard.start + shard.count)
        out["win_row0"] = list(kw["win_row0"])[sl]
        out["win_nrow"] = list(kw["win_nrow"])[sl]
        out["seed_chain"] = seeds[sl, :]
    out["unit_id_base"] = shard.unit_id_base
    return out


def gather_units(local: Optional[np.ndarray], shard: Shard, group=None) -> Optional[np.ndarray]:
    """Gather per-unit arrays ``[local units, ...]`` from every rank onto rank 0 in global unit order
    (the "final gather of draws and forecast densities" - the one exchange of the whole job).
    Uses ``torch.distributed`` (NCCL moves device tensors over NVLink; gloo moves host tensors);
    with world == 1 it is the identity."""
    if shard.world == 1:
        return local
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    shards = [shard_units(shard.n_windows, shard.n_chains, shard.world, r) for r in range(shard.world)]
    # trailing shape from the first non-empty rank
    meta = [None] * shard.world
    dist.all_gather_object(meta, None if local is None else tuple(local.shape[1:]), group=group)
    trail = next(m for m in meta if m is not None)
    mine = torch.zeros((shard.n_units,) + tuple(trail), dtype=torch.float64, device=dev)
    if local is not None and shard.n_units:
        mine.copy_(torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)))
    sizes = [s.n_units for s in shards]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(trail), dtype=torch.float64, device=dev)
    pad[: shard.n_units] = mine
    bufs = [torch.empty_like(pad) for _ in range(shard.world)] if shard.rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if shard.rank != 0:
        return None
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0).cpu().numpy()


def merge_shards(parts: Sequence[Optional[np.ndarray]]) -> np.ndarray:
    """In-process counterpart of ``gather_units``: concatenate the per-shard arrays in rank order."""
    return np.concatenate([p for p in parts if p is not None and len(p)], axis=0)
