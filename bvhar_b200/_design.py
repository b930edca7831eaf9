"""Host-side data preparation for the Gibbs hot path (tiny, once per fit / per window).

Mirrors, function for function, the pieces of the reference that sit just above the sampler:

* ``build_response`` / ``build_design`` / ``build_vhar``  — inst/include/bvhar/src/math/design.h:8-58
* ``minnesota_moments``                                   — design.h:60-98 + bayes/triangular/config.h:124-151
  and triangular.h:438-439 (only the diagonal of the Kronecker precision is used by the sampler)
* ``build_grpmat``                                        — python/src/bvhar/utils/_misc.py:105-118
  (R: R/misc-r.R:346-397)
"""
from __future__ import annotations

import numpy as np


def build_response(y: np.ndarray, var_lag: int, index: int) -> np.ndarray:
    """design.h:8-16 ``build_y0``: rows index-1 ... index-1+n-p-1 of y."""
    y = np.asarray(y, dtype=np.float64)
    n = y.shape[0] - var_lag
    return np.asfortranarray(y[index - 1: index - 1 + n, :])


def build_design(y: np.ndarray, var_lag: int, include_mean: bool = True) -> np.ndarray:
    """design.h:18-33 ``build_x0``: X0 = [Y_p, ..., Y_1, 1]."""
    y = np.asarray(y, dtype=np.float64)
    n = y.shape[0] - var_lag
    cols = [build_response(y, var_lag, var_lag - t) for t in range(var_lag)]
    if include_mean:
        cols.append(np.ones((n, 1)))
    return np.asfortranarray(np.hstack(cols))


def build_vhar(dim: int, week: int, month: int, include_mean: bool = True) -> np.ndarray:
    """design.h:35-58: (3k+1) x (month*k+1) HAR transformation T_HAR (x) I_k, plus the constant."""
    har = np.zeros((3, month))
    har[0, 0] = 1.0
    har[1, :week] = 1.0 / week
    har[2, :month] = 1.0 / month
    out = np.zeros((3 * dim + 1, month * dim + 1))
    out[: 3 * dim, : month * dim] = np.kron(har, np.eye(dim))
    out[3 * dim, month * dim] = 1.0
    if include_mean:
        return np.asfortranarray(out)
    return np.asfortranarray(out[: 3 * dim, : month * dim])


def build_vhar_design(y: np.ndarray, week: int, month: int, include_mean: bool = True) -> np.ndarray:
    """X1 = X0 T_HAR' (forecaster.h:1105-1107, R/vhar-bayes.R:116-118)."""
    x0 = build_design(y, month, include_mean)
    return np.asfortranarray(x0 @ build_vhar(y.shape[1], week, month, include_mean).T)


def build_grpmat(p: int, dim_data: int, minnesota: str = "longrun") -> np.ndarray:
    """_misc.py:105-118: Minnesota group ids of the (p*dim) x dim coefficient block."""
    if minnesota not in ("longrun", "short", "no"):
        raise ValueError(f"Argument ('minnesota')={minnesota} is not valid: Choose between ['longrun', 'short', 'no']")
    if minnesota == "no":
        return np.full((p * dim_data, dim_data), 1, dtype=np.int32)
    res = [np.identity(dim_data) + 1]
    for i in range(1, p):
        if minnesota == "longrun":
            res.append(np.identity(dim_data) + (i * 2 + 1))
        else:
            res.append(np.full((dim_data, dim_data), i + 2))
    return np.vstack(res).astype(np.int32)


def _build_ydummy(p, sigma, lam, daily, weekly, monthly):
    """design.h:60-78 with include_mean = false."""
    dim = sigma.size
    res = np.zeros((dim * p + dim, dim))
    res[:dim, :dim][np.diag_indices(dim)] = daily * sigma / lam
    if p > 1:
        res[dim:2 * dim, :dim][np.diag_indices(dim)] = weekly * sigma / lam
        res[2 * dim:3 * dim, :dim][np.diag_indices(dim)] = monthly * sigma / lam
    res[dim * p: dim * p + dim, :dim][np.diag_indices(dim)] = sigma
    return res


def _build_xdummy(lag_seq, lam, sigma, eps):
    """design.h:80-98 with include_mean = false (eps only touches the dropped constant row)."""
    dim = sigma.size
    p = lag_seq.size
    res = np.zeros((dim * p + dim, dim * p))
    res[: dim * p, : dim * p] = np.kron(np.diag(lag_seq), np.diag(sigma / lam))
    return res


def minnesota_moments(priors: dict, dim: int, nrow: int, hierarchical: bool = False):
    """Prior mean (nrow x dim) and the diagonal prior precision in alpha order (nrow*dim,).

    config.h:124-151 (Minnesota, uses ``lambda``) and config.h:182-208 (hierarchical: lambda = 1).
    """
    lag = int(priors["p"])
    sigma = np.asarray(priors["sigma"], dtype=np.float64).reshape(-1)
    if sigma.size != dim or lag * dim != nrow:
        raise ValueError("Minnesota prior: 'sigma' / 'p' do not match the coefficient matrix")
    lam = 1.0 if hierarchical else float(priors["lambda"])
    eps = float(priors["eps"])
    if "delta" in priors:
        daily = np.asarray(priors["delta"], dtype=np.float64).reshape(-1)
        weekly = np.zeros(dim); monthly = np.zeros(dim)
    else:
        daily = np.asarray(priors["daily"], dtype=np.float64).reshape(-1)
        weekly = np.asarray(priors["weekly"], dtype=np.float64).reshape(-1)
        monthly = np.asarray(priors["monthly"], dtype=np.float64).reshape(-1)
    yd = _build_ydummy(lag, sigma, lam, daily, weekly, monthly)
    xd = _build_xdummy(np.arange(1, lag + 1, dtype=np.float64), lam, sigma, eps)
    prec = xd.T @ xd
    chol = np.linalg.cholesky(prec)
    rhs = xd.T @ yd
    mean = np.linalg.solve(chol.T, np.linalg.solve(chol, rhs))
    alpha_prec = np.kron(1.0 / sigma, np.diag(prec))  # diag(kron(diag(1/sigma), prec)), triangular.h:439
    return np.asfortranarray(mean), alpha_prec


def build_dummies(p: int, sigma, lam: float, daily, weekly, monthly, eps: float, include_mean: bool = True):
    """``build_ydummy`` / ``build_xdummy`` (design.h:60-98) including the constant-term row / column."""
    sigma = np.asarray(sigma, dtype=np.float64).reshape(-1)
    dim = sigma.size
    yd = np.zeros((dim * p + dim + 1, dim))
    yd[:dim * p + dim] = _build_ydummy(p, sigma, lam, np.asarray(daily, dtype=np.float64), np.asarray(weekly, dtype=np.float64),
                                      np.asarray(monthly, dtype=np.float64))
    xd = np.zeros((dim * p + dim + 1, dim * p + 1))
    xd[:dim * p + dim, :dim * p] = _build_xdummy(np.arange(1, p + 1, dtype=np.float64), lam, sigma, eps)
    xd[dim * p + dim, dim * p] = eps
    if include_mean:
        return xd, yd
    return xd[:dim * p + dim, :dim * p], yd[:dim * p + dim]


def mniw_posterior(x, y, p: int, sigma, lam: float, delta, eps: float = 1e-4, weekly=None, monthly=None,
                   include_mean: bool = True):
    """Closed-form Minnesota (matrix-normal inverse-Wishart) posterior on dummy-augmented data:
    ``Minnesota::estimateCoef`` / ``estimateCov`` (bayes/mniw/minnesota.h:153-186).

    prec = X*'X*, coef = prec^-1 X*'Y*, IW scale = (Y* - X* coef)'(Y* - X* coef),
    IW shape = n_dummy - dim_design + 2 + n, with X* = [X; Xd], Y* = [Y; Yd].
    Host-side and one-off: a deterministic parity target, not a kernel."""
    x = np.asarray(x, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    dim = y.shape[1]
    zeros = np.zeros(dim)
    xd, yd = build_dummies(p, sigma, lam, delta, zeros if weekly is None else weekly, zeros if monthly is None else monthly,
                           eps, include_mean)
    xs = np.vstack([x, xd]); ys = np.vstack([y, yd])
    prec = xs.T @ xs
    chol = np.linalg.cholesky(prec)
    coef = np.linalg.solve(chol.T, np.linalg.solve(chol, xs.T @ ys))
    res = ys - xs @ coef
    return {"coef": coef, "prec": prec, "scale": res.T @ res, "shape": xd.shape[0] - x.shape[1] + 2 + x.shape[0],
            "x_dummy": xd, "y_dummy": yd}
