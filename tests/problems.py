"""Seeded synthetic problems shared by the oracle and the CUDA parity tests."""
from __future__ import annotations

import numpy as np

from bvhar_b200 import _design as dz
from bvhar_b200 import _spec as sp
from bvhar_b200._abi import PackedFit, REC_ALL

PRIORS = ["Minnesota", "SSVS", "Horseshoe", "HMN", "NG", "DL", "GDP"]


def simulate_series(T, dim, seed):
    rng = np.random.default_rng(seed)
    y = np.zeros((T + 50, dim))
    A = 0.4 * np.eye(dim) + rng.uniform(-0.05, 0.05, (dim, dim))
    L = np.eye(dim) + np.tril(rng.normal(0, 0.2, (dim, dim)), -1)
    h = np.zeros(dim)
    for t in range(1, T + 50):
        h = 0.95 * h + rng.normal(0, 0.15, dim)
        e = np.linalg.solve(L, np.exp(h / 2) * rng.normal(size=dim))
        y[t] = 0.1 + y[t - 1] @ A.T + 0.5 * e
    return y[50:]


def make_problem(prior="DL", sv=True, dim=3, T=40, lag=2, vhar=False, week=2, month=4, chains=2, iters=6, burn=2,
                 thin=1, seed=7, include_mean=True, ggl=True, minnesota=True, record_mask=REC_ALL, windows=None,
                 y=None):
    """Returns (PackedFit kwargs dict, meta).  `windows` = list of (row0, nrow) on the full design."""
    rng = np.random.default_rng(seed)
    if y is None:
        y = simulate_series(T, dim, seed + 1000)
    dim = y.shape[1]
    if vhar:
        p = 3
        x = dz.build_vhar_design(y, week, month, include_mean)
        resp = dz.build_response(y, month, month + 1)
        grp = dz.build_grpmat(3, dim, "longrun" if minnesota else "no")
        own, cross = ([2, 4, 6], [1, 3, 5]) if minnesota else ([1], [2])
    else:
        p = lag
        x = dz.build_design(y, lag, include_mean)
        resp = dz.build_response(y, lag, lag + 1)
        grp = dz.build_grpmat(lag, dim, "short" if minnesota else "no")
        own, cross = (list(range(2, 2 * p, 2)) or [2], list(range(1, 2 * p, 2))) if minnesota else ([2], [2])
        if minnesota:  # python/src/bvhar/model/_bayes.py:309-310
            own, cross = list(np.arange(2, 2 * p, 2)), list(np.arange(1, 2 * p, 2))
    grp_id = np.array(sorted(set(grp.reshape(-1).tolist())), dtype=np.int32)
    grp_id = np.array(list(dict.fromkeys(grp.flatten(order="F").tolist())), dtype=np.int32)  # pd.unique order
    cov = sp.SvConfig() if sv else sp.LdltConfig()
    cov.update(dim)
    icpt = sp.InterceptConfig()
    icpt.update(dim)
    cfg = {"Minnesota": sp.MinnesotaConfig(is_long=vhar), "SSVS": sp.SsvsConfig(), "Horseshoe": sp.HorseshoeConfig(),
           "HMN": sp.MinnesotaConfig(lam=sp.LambdaConfig(), is_long=vhar), "NG": sp.NgConfig(), "DL": sp.DlConfig(),
           "GDP": sp.GdpConfig()}[prior]
    if prior == "SSVS":
        cfg.update(grp_id, np.array(own), np.array(cross))
    elif prior in ("Minnesota", "HMN"):
        cfg.update(y, p, dim)
    n_obs = resp.shape[0] if windows is None else windows[0][1]
    inits = sp.make_inits(prior, sv, dim, x.shape[1], n_obs, len(grp_id), chains, rng, include_mean=include_mean)
    n_win = 1 if windows is None else len(windows)
    seeds = rng.integers(1, 2**31 - 1, size=(n_win, chains))
    kw = dict(num_chains=chains, num_iter=iters, num_burn=burn, thin=thin, x=x, y=resp, param_cov=cov.to_dict(),
              param_prior=cfg.to_dict(), param_intercept=icpt.to_dict(), param_init=inits,
              prior_type=sp.PRIOR_TYPE[cfg.prior], grp_id=grp_id, own_id=np.array(own, dtype=np.int32),
              cross_id=np.array(cross, dtype=np.int32), grp_mat=grp, include_mean=include_mean, seed_chain=seeds,
              is_group=ggl, record_mask=record_mask)
    if windows is not None:
        kw["win_row0"] = [w[0] for w in windows]
        kw["win_nrow"] = [w[1] for w in windows]
    return kw, dict(y=y, x=x, resp=resp, dim=dim, p=p, grp=grp, grp_id=grp_id)


def pack(**kwargs):
    kw, meta = make_problem(**kwargs)
    return PackedFit(**kw), meta
