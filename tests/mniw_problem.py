"""A small Minnesota posterior (closed form, minnesota.h:153-186) for the conjugate-sampler tests."""
import numpy as np

from bvhar_b200 import _design


def posterior(dim=3, p=2, T=80, seed=5, include_mean=True):
    rng = np.random.default_rng(seed)
    y = np.cumsum(rng.normal(size=(T, dim)), axis=0) * 0.1 + rng.normal(size=(T, dim))
    x = _design.build_design(y, p, include_mean)
    y0 = _design.build_response(y, p, p + 1)
    sigma = y.std(axis=0, ddof=1)
    post = _design.mniw_posterior(x, y0, p, sigma, 0.2, np.ones(dim), 1e-4, include_mean=include_mean)
    return x, y0, post


def mh_inits(dim, chains, rng):
    """(lambda, psi) modes, a positive-definite 'hessian' and the scale of the proposal, as optim() would return them."""
    np_ = 1 + dim
    par = np.column_stack([rng.uniform(.15, .3, chains)] + [rng.uniform(.8, 1.5, chains) for _ in range(dim)])
    hess = []
    for _ in range(chains):
        a = rng.normal(size=(np_ + 3, np_))
        hess.append(a.T @ a * 40.0 + np.eye(np_) * 5.0)
    return par, np.stack(hess), np.full(chains, 0.05)
