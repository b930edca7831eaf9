"""Conjugate Minnesota samplers on the device.

Mirrors of the two exported estimators of the reference's ``src/mniw.cpp`` - same names, argument order and result
layout (one dict of records per chain):

* :func:`estimate_mniw`    - ``estimate_mniw`` (src/mniw.cpp:130-190): ``McmcMniw`` (bayes/mniw/minnesota.h:241-278),
  ``num_iter`` draws of ``sim_mn_iw`` from a fixed matrix-normal inverse-Wishart posterior;
* :func:`estimate_bvar_mh` - ``estimate_bvar_mh`` (src/mniw.cpp:46-110): ``MhMinnesota`` (minnesota.h:413-510), the same
  draws plus the random-walk Metropolis step on (lambda, psi).

Both go through ``bvhar_cuda_mniw_run`` (include/bvhar_cuda.h): every (chain, kept iteration) is one independent CTA.
The closed-form posterior moments (``Minnesota::estimateCoef`` / ``estimateCov``, minnesota.h:176-187) are the host-side
``_design.mniw_posterior``-style products, as in round 1.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence

import numpy as np

from . import _abi


def _run(num_chains, num_iter, num_burn, thin, mn_mean, mn_prec, iw_scale, iw_shape, seed_chain, mh=None, device=0, unit_id_base=0):
    mean = np.asfortranarray(mn_mean, dtype=np.float64)
    prec = np.asfortranarray(mn_prec, dtype=np.float64)
    scale = np.asfortranarray(iw_scale, dtype=np.float64)
    d, k = mean.shape
    if prec.shape != (d, d) or scale.shape != (k, k):
        raise ValueError("mn_prec must be dim_design x dim_design and iw_scale dim x dim")
    seeds = np.ascontiguousarray(np.asarray(seed_chain).reshape(-1), dtype=np.uint32)
    if seeds.size != num_chains:
        raise ValueError("seed_chain needs one seed per chain")
    thin = max(int(thin), 1)
    rows = (num_iter - num_burn + thin - 1) // thin
    D = _abi.MniwDesc()
    D.device = device; D.num_chains = num_chains; D.num_iter = num_iter; D.num_burn = num_burn; D.thin = thin
    D.dim = k; D.dim_design = d
    D.mn_mean = mean.ctypes.data_as(_abi.c_dp); D.mn_prec = prec.ctypes.data_as(_abi.c_dp); D.iw_scale = scale.ctypes.data_as(_abi.c_dp)
    D.iw_shape = float(iw_shape)
    D.seed_chain = seeds.ctypes.data_as(_abi.c_up); D.unit_id_base = unit_id_base
    keep = [mean, prec, scale, seeds]
    alpha = np.empty((num_chains, d * k, rows)); sigma = np.empty((num_chains, k * k, rows))
    lam = psi = acc = None
    if mh is not None:
        par = np.ascontiguousarray(mh["par"], dtype=np.float64).reshape(num_chains, 1 + k)
        hess = np.ascontiguousarray(np.stack([np.asfortranarray(h, dtype=np.float64).T for h in mh["hessian"]]))  # col-major blocks
        sv = np.ascontiguousarray(mh["scale_variance"], dtype=np.float64).reshape(num_chains)
        keep += [par, hess, sv]
        D.mh = 1
        D.gam_shape, D.gam_rate = float(mh["lambda_param"][0]), float(mh["lambda_param"][1])
        D.invgam_shape, D.invgam_scl = float(mh["psi_param"][0]), float(mh["psi_param"][1])
        D.init_par = par.ctypes.data_as(_abi.c_dp); D.init_hess = hess.ctypes.data_as(_abi.c_dp)
        D.init_scale_variance = sv.ctypes.data_as(_abi.c_dp)
        lam = np.empty((num_chains, rows)); psi = np.empty((num_chains, k, rows)); acc = np.empty((num_chains, rows), dtype=np.uint8)
    L = _abi.load_library()
    _abi.check(L.bvhar_cuda_mniw_run(C.byref(D), alpha.ctypes.data_as(_abi.c_dp), sigma.ctypes.data_as(_abi.c_dp),
                                     lam.ctypes.data_as(_abi.c_dp) if lam is not None else None,
                                     psi.ctypes.data_as(_abi.c_dp) if psi is not None else None,
                                     acc.ctypes.data_as(C.POINTER(C.c_uint8)) if acc is not None else None))
    res: List[Dict[str, np.ndarray]] = []
    for c in range(num_chains):  # [cols][rows] blocks are the column-major rows x cols record matrices of the reference
        out = {}
        if mh is not None:
            out["lambda_record"] = lam[c].copy()
            out["psi_record"] = psi[c].T
        out["alpha_record"] = alpha[c].T
        out["sigma_record"] = sigma[c].T
        if mh is not None:
            out["accept_record"] = acc[c].astype(bool)
        res.append(out)
    return res


def estimate_mniw(num_chains: int, num_iter: int, num_burn: int, thin: int, mn_mean, mn_prec, iw_scale, iw_shape: float,
                  seed_chain: Sequence[int], display_progress: bool = False, nthreads: int = 1, device: int = 0):
    """``estimate_mniw`` (src/mniw.cpp:130-190).  Returns one ``{"alpha_record", "sigma_record"}`` dict per chain:
    ``alpha_record`` rows are ``vec`` of the ``dim_design x dim`` coefficient draw (minnesota.h:128), ``sigma_record`` rows
    ``vec(Sigma)``.  ``display_progress`` / ``nthreads`` are accepted for signature parity (every draw is one CTA)."""
    return _run(num_chains, num_iter, num_burn, thin, mn_mean, mn_prec, iw_scale, iw_shape, seed_chain, device=device)


def estimate_bvar_mh(num_chains: int, num_iter: int, num_burn: int, thin: int, x, y, x_dummy, y_dummy, param_prior: dict,
                     param_init: Sequence[dict], seed_chain: Sequence[int], display_progress: bool = False, nthreads: int = 1,
                     device: int = 0):
    """``estimate_bvar_mh`` (src/mniw.cpp:46-110).  ``param_prior = {"lambda": {"param": (shape, rate)}, "sigma":
    {"param": (shape, scale)}}`` (MhMinnSpec, minnesota.h:89-105); ``param_init[i] = {"par": (lambda, psi...), "hessian",
    "scale_variance"}`` (MhMinnInits, minnesota.h:71-87).  Returns ``lambda_record, psi_record, alpha_record,
    sigma_record, accept_record`` per chain (minnesota.h:470-478)."""
    x = np.asarray(x, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    xd = np.asarray(x_dummy, dtype=np.float64); yd = np.asarray(y_dummy, dtype=np.float64)
    # Minnesota ctor + computePosterior (minnesota.h:153-186, 436-439)
    xs = np.vstack([x, xd]); ys = np.vstack([y, yd])
    prec = xs.T @ xs
    chol = np.linalg.cholesky(prec)
    coef = np.linalg.solve(chol.T, np.linalg.solve(chol, xs.T @ ys))
    resid = ys - xs @ coef
    scale = resid.T @ resid
    shape = xd.shape[0] - x.shape[1] + 2 + x.shape[0]  # prior_shape + num_design
    mh = {"par": [np.asarray(p["par"], dtype=np.float64) for p in param_init],
          "hessian": [np.asarray(p["hessian"], dtype=np.float64) for p in param_init],
          "scale_variance": [float(p["scale_variance"]) for p in param_init],
          "lambda_param": np.asarray(param_prior["lambda"]["param"], dtype=np.float64),
          "psi_param": np.asarray(param_prior["sigma"]["param"], dtype=np.float64)}
    return _run(num_chains, num_iter, num_burn, thin, coef, prec, scale, shape, seed_chain, mh=mh, device=device)
