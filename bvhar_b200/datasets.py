"""Input data for the BASELINE.json configurations.

* ``load_vix()`` — the packaged CBOE ETF volatility index panel (905 x 9), the data file shipped by the
  reference as ``etf_vix`` (python/src/bvhar/datasets/data/etf_vix.csv; R: data/etf_vix.rda).
* ``simulate_vhar_sv`` / ``simulate_var_sv`` — the seeded synthetic series of SURVEY.md section 8(d)
  for the configurations that have no packaged data (C3, C4, C5); standardised to unit variance.
"""
from __future__ import annotations

import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_vix() -> np.ndarray:
    return np.asfortranarray(np.loadtxt(os.path.join(_DATA, "etf_vix.csv"), delimiter=",", skiprows=1))


def _sv_innovations(T, dim, rng, sig_h=0.1, chol_sd=0.1):
    L = np.eye(dim) + np.tril(rng.normal(0, chol_sd, (dim, dim)), -1)
    Linv = np.linalg.inv(L)
    h = np.cumsum(rng.normal(0, sig_h, (T, dim)), axis=0)
    h -= h.mean(axis=0)
    return (np.exp(h / 2) * rng.normal(size=(T, dim))) @ Linv.T


def simulate_vhar_sv(T: int = 1022, dim: int = 20, week: int = 5, month: int = 22, seed: int = 20240001) -> np.ndarray:
    """Stable VHAR(week, month) with own-lag effects .3 / .2 / .1, cross effects U(-.02, .02), SV innovations."""
    rng = np.random.default_rng(seed)
    phi = [np.diag(np.full(dim, v)) + rng.uniform(-.02, .02, (dim, dim)) * (1 - np.eye(dim)) for v in (.3, .2, .1)]
    burn = 200
    e = 0.3 * _sv_innovations(T + burn, dim, rng)
    y = np.zeros((T + burn, dim))
    for t in range(month, T + burn):
        d = y[t - 1]
        w = y[t - week:t].mean(axis=0)
        m = y[t - month:t].mean(axis=0)
        y[t] = phi[0] @ d + phi[1] @ w + phi[2] @ m + e[t]
    y = y[burn:]
    return np.asfortranarray((y - y.mean(axis=0)) / y.std(axis=0))


def simulate_var_sv(T: int, dim: int, lag: int = 4, seed: int = 20240002, own=(.4, .2, .1, .05), cross_frac=.05,
                    cross_sd=.05) -> np.ndarray:
    """Stable VAR(lag) with decaying own-lag effects and a sparse set of N(0, cross_sd^2) cross effects."""
    rng = np.random.default_rng(seed)
    A = []
    for l in range(lag):
        a = np.diag(np.full(dim, own[l] if l < len(own) else 0.0))
        mask = (rng.uniform(size=(dim, dim)) < cross_frac) & ~np.eye(dim, dtype=bool)
        a = a + mask * rng.normal(0, cross_sd, (dim, dim))
        A.append(a)
    burn = 200
    e = 0.3 * _sv_innovations(T + burn, dim, rng)
    y = np.zeros((T + burn, dim))
    for t in range(lag, T + burn):
        y[t] = sum(A[l] @ y[t - 1 - l] for l in range(lag)) + e[t]
    y = y[burn:]
    return np.asfortranarray((y - y.mean(axis=0)) / y.std(axis=0))
