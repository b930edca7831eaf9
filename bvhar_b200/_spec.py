"""Prior / covariance configuration objects of the Bayesian VAR / VHAR front end.

Same class names, constructor arguments, defaults and ``to_dict()`` keys as the reference's Python
configuration layer (python/src/bvhar/model/_spec.py:4-384; R equivalents ``set_ldlt``, ``set_sv``,
``set_intercept``, ``set_ssvs``, ``set_horseshoe``, ``set_bvar``/``set_bvhar``, ``set_lambda``,
``set_dl``, ``set_ng``, ``set_gdp`` in R/hyperparam.R:64-788).  The dictionaries produced here are
the ``param_cov`` / ``param_prior`` / ``param_intercept`` lists whose keys the sampler parses
(inst/include/bvhar/src/bayes/triangular/config.h:71-74, 100-101, 124-143, 180, 243-248, 298-301,
327, 352).
"""
from __future__ import annotations

from math import sqrt

import numpy as np

_NUM = (int, float, np.number)


def _as_vec(value, member):
    if isinstance(value, _NUM):
        return np.array([float(value)])
    if isinstance(value, (list, tuple, np.ndarray)):
        if len(value) == 0:
            raise ValueError(f"'{member}' cannot be empty.")
        return np.asarray(value, dtype=np.float64)
    raise TypeError(f"'{member}' should be a number or a numeric array.")


def _num_or_array(value, member, n_size=None):
    if isinstance(value, _NUM):
        return value
    if isinstance(value, (list, tuple, np.ndarray)):
        if len(value) == 0:
            raise ValueError(f"'{member}' cannot be empty.")
        if n_size is not None and len(value) != n_size:
            raise ValueError(f"'{member}' length must be {n_size}.")
        arr = np.asarray(value)
        if arr.ndim > 2:
            raise ValueError(f"'{member} has wrong dim = {arr.ndim}.")
        return arr
    raise TypeError(f"'{member}' should be a number or a numeric array.")


class LdltConfig:
    """Inverse-gamma prior on the diagonal of the LDLT factor (homoskedastic model)."""

    def __init__(self, ig_shape=3, ig_scale=.01):
        self.process = "Homoskedastic"
        self.prior = "Cholesky"
        self.shape = _as_vec(ig_shape, "ig_shape")
        self.scale = _as_vec(ig_scale, "ig_scale")

    def update(self, n_dim: int):
        if len(self.shape) == 1:
            self.shape = np.repeat(self.shape, n_dim)
        if len(self.scale) == 1:
            self.scale = np.repeat(self.scale, n_dim)

    def to_dict(self):
        return {"shape": self.shape, "scale": self.scale}


class SvConfig(LdltConfig):
    """Stochastic volatility: IG prior on the state variances, N(mean, prec^-1) on h_0."""

    def __init__(self, ig_shape=3, ig_scale=.01, initial_mean=1, initial_prec=.1):
        super().__init__(ig_shape, ig_scale)
        self.process = "SV"
        self.initial_mean = _as_vec(initial_mean, "initial_mean")
        if isinstance(initial_prec, _NUM):
            self.initial_prec = float(initial_prec) * np.identity(1)
        else:
            self.initial_prec = np.asarray(initial_prec, dtype=np.float64)
            if self.initial_prec.ndim != 2:
                raise ValueError("'initial_prec' should be 2-dim.")

    def update(self, n_dim: int):
        super().update(n_dim)
        if len(self.initial_mean) == 1:
            self.initial_mean = np.repeat(self.initial_mean, n_dim)
        if self.initial_prec.size == 1:
            self.initial_prec = float(self.initial_prec.reshape(-1)[0]) * np.identity(n_dim)

    def to_dict(self):
        return {"shape": self.shape, "scale": self.scale,
                "initial_mean": self.initial_mean, "initial_prec": self.initial_prec}


class InterceptConfig:
    """Normal prior for the constant term."""

    def __init__(self, mean=0, sd=.1):
        self.process = "Intercept"
        self.prior = "Normal"
        self.mean_non = _as_vec(mean, "mean")
        if not isinstance(sd, _NUM):
            raise ValueError("'sd_non' should be a number.")
        self.sd_non = float(sd)

    def update(self, n_dim: int):
        if len(self.mean_non) == 1:
            self.mean_non = np.repeat(self.mean_non, n_dim)

    def to_dict(self):
        return {"mean_non": self.mean_non, "sd_non": self.sd_non}


class _BayesConfig:
    def __init__(self, prior):
        self.prior = prior

    def update(self, *args, **kwargs):
        pass

    def to_dict(self):
        return {}


class SsvsConfig(_BayesConfig):
    def __init__(self, coef_spike_grid=100, coef_slab_shape=.01, coef_slab_scl=.01, coef_s1=[1, 1], coef_s2=[1, 1],
                 chol_spike_grid=100, chol_slab_shape=.01, chol_slab_scl=.01, chol_s1=1, chol_s2=1):
        super().__init__("SSVS")
        self.coef_spike_grid = _num_or_array(coef_spike_grid, "coef_spike_grid")
        self.coef_slab_shape = _num_or_array(coef_slab_shape, "coef_slab_shape")
        self.coef_slab_scl = _num_or_array(coef_slab_scl, "coef_slab_scl")
        self.coef_s1 = _num_or_array(coef_s1, "coef_s1", 2)
        self.coef_s2 = _num_or_array(coef_s2, "coef_s2", 2)
        self.chol_spike_grid = _num_or_array(chol_spike_grid, "chol_spike_grid")
        self.chol_slab_shape = _num_or_array(chol_slab_shape, "chol_slab_shape")
        self.chol_slab_scl = _num_or_array(chol_slab_scl, "chol_slab_scl")
        self.chol_s1 = _num_or_array(chol_s1, "chol_s1")
        self.chol_s2 = _num_or_array(chol_s2, "chol_s2")

    def update(self, grp_id, own_id, cross_id):
        # (own, cross) pair -> one Beta hyper-parameter per Minnesota group
        for name in ("coef_s1", "coef_s2"):
            val = getattr(self, name)
            if len(val) == 2 and len(grp_id) != 2:
                full = np.zeros(len(grp_id))
                full[np.isin(grp_id, own_id)] = val[0]
                full[np.isin(grp_id, cross_id)] = val[1]
                setattr(self, name, full)
            elif len(val) == 2:
                full = np.zeros(2)
                full[np.isin(grp_id, own_id)] = val[0]
                full[np.isin(grp_id, cross_id)] = val[1]
                setattr(self, name, full)

    def to_dict(self):
        return {"coef_grid": self.coef_spike_grid, "coef_slab_shape": self.coef_slab_shape,
                "coef_slab_scl": self.coef_slab_scl, "coef_s1": self.coef_s1, "coef_s2": self.coef_s2,
                "chol_grid": self.chol_spike_grid, "chol_slab_shape": self.chol_slab_shape,
                "chol_slab_scl": self.chol_slab_scl, "chol_s1": self.chol_s1, "chol_s2": self.chol_s2}


class HorseshoeConfig(_BayesConfig):
    def __init__(self):
        super().__init__("Horseshoe")


class LambdaConfig:
    r"""Gamma hyper-prior on the Minnesota tightness :math:`\lambda` (hierarchical Minnesota)."""

    def __init__(self, shape=.01, rate=.01, grid_size=100, eps=1e-04):
        self.shape_ = float(shape)
        self.rate_ = float(rate)
        self.grid_size = int(grid_size)

    def update(self, mode, sd):
        self.shape_ = (2 + mode ** 2 / (sd ** 2) + sqrt((2 + mode ** 2 / (sd ** 2)) ** 2 - 4)) / 2
        self.rate_ = sqrt(self.shape_) / sd


class MinnesotaConfig(_BayesConfig):
    def __init__(self, sig=None, lam=.1, delt=None, is_long=False, eps=1e-04):
        super().__init__("Minnesota")
        self.sig = None if sig is None else _num_or_array(sig, "sig")
        self.lam = lam
        if isinstance(lam, LambdaConfig):
            self.lam = [lam.shape_, lam.rate_, lam.grid_size]
            self.prior = "HMN"
        self.delt = None if delt is None else _num_or_array(delt, "delt")
        self.eps = _num_or_array(eps, "eps")
        self.p = None
        self.is_long = bool(is_long)
        if is_long:
            self.weekly = None
            self.monthly = None

    def update(self, y, p, n_dim: int):
        self.p = p
        if self.sig is None:
            self.sig = np.std(y, axis=0)
        if self.delt is None:
            self.delt = np.repeat(0, n_dim)
        if self.is_long:
            self.weekly = np.repeat(0, n_dim)
            self.monthly = np.repeat(0, n_dim)

    def to_dict(self):
        out = {"p": self.p, "sigma": self.sig, "eps": self.eps}
        if isinstance(self.lam, list):
            out.update({"shape": self.lam[0], "rate": self.lam[1], "grid_size": self.lam[2]})
        else:
            out["lambda"] = self.lam
        if self.is_long:
            out.update({"daily": self.delt, "weekly": self.weekly, "monthly": self.monthly})
        else:
            out["delta"] = self.delt
        return out


class DlConfig(_BayesConfig):
    def __init__(self, dir_grid: int = 100, shape=.01, scale=.01):
        super().__init__("DL")
        self.grid_size = _num_or_array(dir_grid, "dir_grid")
        self.shape = _num_or_array(shape, "shape")
        self.scale = _num_or_array(scale, "scale")

    def to_dict(self):
        return {"grid_size": self.grid_size, "shape": self.shape, "scale": self.scale}


class NgConfig(_BayesConfig):
    def __init__(self, shape_sd=.01, group_shape=.01, group_scale=.01, global_shape=.01, global_scale=.01,
                 contem_global_shape=.01, contem_global_scale=.01):
        super().__init__("NG")
        self.shape_sd = _num_or_array(shape_sd, "shape_sd")
        self.group_shape = _num_or_array(group_shape, "group_shape")
        self.group_scale = _num_or_array(group_scale, "group_scale")
        self.global_shape = _num_or_array(global_shape, "global_shape")
        self.global_scale = _num_or_array(global_scale, "global_scale")
        self.contem_global_shape = _num_or_array(contem_global_shape, "contem_global_shape")
        self.contem_global_scale = _num_or_array(contem_global_scale, "contem_global_scale")

    def to_dict(self):
        return {"shape_sd": self.shape_sd, "group_shape": self.group_shape, "group_scale": self.group_scale,
                "global_shape": self.global_shape, "global_scale": self.global_scale,
                "contem_global_shape": self.contem_global_shape, "contem_global_scale": self.contem_global_scale}


class GdpConfig(_BayesConfig):
    def __init__(self, shape_grid: int = 100, rate_grid: int = 100):
        super().__init__("GDP")
        self.grid_shape = _num_or_array(shape_grid, "shape_grid")
        self.grid_rate = _num_or_array(rate_grid, "rate_grid")

    def to_dict(self):
        return {"grid_shape": self.grid_shape, "grid_rate": self.grid_rate}


PRIOR_TYPE = {"Minnesota": 1, "SSVS": 2, "Horseshoe": 3, "HMN": 4, "NG": 5, "DL": 6, "GDP": 7}


def make_inits(prior: str, sv: bool, dim: int, n_design: int, n_obs: int, n_grp: int, n_chain: int, rng, include_mean=None):
    """Random initial values per chain, drawn as the reference front ends draw them
    (python/src/bvhar/model/_bayes.py:58-170; R/var-bayes.R:164-172, 271-413, 440-463):
    U(-1,1) coefficients, exp(U(-1,0)) contemporaneous terms, exp(U(-1,1)) scales.

    ``rng`` is a ``numpy.random.Generator``; ``n_design`` = rows of the coefficient matrix
    (constant included), ``n_obs`` = rows of the response.
    """
    n_eta = dim * (dim - 1) // 2
    has_const = (n_design % dim != 0) if include_mean is None else bool(include_mean)
    n_alpha = dim * (n_design - (1 if has_const else 0))
    inits = []
    for _ in range(n_chain):
        ini = {
            "init_coef": np.asfortranarray(rng.uniform(-1, 1, (n_design, dim))),
            "init_contem": np.exp(rng.uniform(-1, 0, n_eta)),
        }
        if sv:
            ini.update({
                "lvol_init": rng.uniform(-1, 1, dim),
                "lvol": np.asfortranarray(np.exp(rng.uniform(-1, 1, (n_obs, dim)))),
                "lvol_sig": np.exp(rng.uniform(-1, 1, dim)),  # R/var-bayes.R:446-448 (length dim)
            })
        else:
            ini["init_diag"] = np.exp(rng.uniform(-1, 1, dim))
        e = lambda size=None: np.exp(rng.uniform(-1, 1, size))
        if prior == "SSVS":
            logistic = lambda x: np.exp(x) / (1 + np.exp(x))
            ini.update({
                "init_coef_dummy": rng.binomial(1, .5, n_alpha).astype(np.float64),
                "coef_mixture": logistic(rng.uniform(-1, 1, n_grp)),
                "coef_slab": e(n_alpha),
                "chol_mixture": logistic(rng.uniform(-1, 1, n_eta)),
                "contem_slab": e(n_eta),
                "coef_spike_scl": rng.uniform(0, 1),
                "chol_spike_scl": rng.uniform(0, 1),
            })
        elif prior in ("Horseshoe", "NG", "DL"):
            ini.update({
                "local_sparsity": e(n_alpha), "global_sparsity": float(e()),
                "contem_local_sparsity": e(n_eta), "contem_global_sparsity": np.array([float(e())]),
            })
            if prior in ("Horseshoe", "NG"):
                ini["group_sparsity"] = e(n_grp)
            if prior == "NG":
                ini["local_shape"] = rng.uniform(0, 1, n_grp)
                ini["contem_shape"] = rng.uniform(0, 1)
        elif prior == "HMN":
            ini.update({"own_lambda": rng.uniform(0, 1), "cross_lambda": rng.uniform(0, 1),
                        "contem_lambda": rng.uniform(0, 1)})
        elif prior == "GDP":
            ini.update({
                "local_sparsity": e(n_alpha), "group_rate": e(n_grp),
                "contem_local_sparsity": e(n_eta), "contem_rate": e(n_eta),
                "gamma_shape": rng.uniform(0, 1), "gamma_rate": rng.uniform(0, 1),
                "contem_gamma_shape": rng.uniform(0, 1), "contem_gamma_rate": rng.uniform(0, 1),
            })
        inits.append(ini)
    return inits
