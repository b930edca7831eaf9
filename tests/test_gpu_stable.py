"""bvhar_cuda_is_stable (the forecasters' stability filter, draw.h:1265-1295, config.h:884-914) against the companion-matrix
eigenvalues computed on the host with LAPACK."""
import numpy as np
import pytest

from bvhar_b200._design import build_vhar
from bvhar_b200._engine import _stable_rows

pytestmark = pytest.mark.gpu


def _radius(alpha, dim, nrow, har=None):
    A = alpha[: nrow * dim].reshape(nrow, dim, order="F")
    if har is not None:
        A = har.T @ A
    m = A.shape[0]
    comp = np.zeros((m, m))
    comp[:dim, :] = A.T
    comp[dim:, : m - dim] = np.eye(m - dim)
    return np.max(np.abs(np.linalg.eigvals(comp)))


def _draws(rng, n, dim, nrow, scales):
    """Random coefficient draws whose companion spectral radius sweeps through 1 (plus a constant column, ignored)."""
    out = np.empty((n, nrow * dim + dim))
    for i in range(n):
        A = rng.normal(size=(nrow, dim)) * 0.15
        A[:dim] += np.eye(dim) * rng.uniform(-0.9, 0.9)
        out[i, : nrow * dim] = (A * scales[i % len(scales)]).reshape(-1, order="F")
        out[i, nrow * dim:] = rng.normal(size=dim)
    return out


@pytest.mark.parametrize("dim,lag", [(1, 1), (2, 3), (3, 5), (9, 5), (20, 2)])
def test_var_draws_match_eigenvalues(dim, lag):
    rng = np.random.default_rng(100 * dim + lag)
    nrow = dim * lag
    coef = _draws(rng, 96, dim, nrow, [0.3, 0.8, 1.0, 1.3, 2.0])
    rho = np.array([_radius(c, dim, nrow) for c in coef])
    far = np.abs(rho - 1) > 1e-9
    got = _stable_rows(coef, dim, nrow, None)
    assert far.all() and 5 < (rho < 1).sum() < 91
    assert np.array_equal(got, rho < 1)


def test_radius_rescaled_to_the_boundary():
    """Draws rescaled (lag-wise, A_l -> A_l / c^l scales every companion eigenvalue by 1 / c) to rho = 1 -+ 1e-6,
    complex dominant pairs included."""
    rng = np.random.default_rng(7)
    dim, lag = 4, 3
    nrow = dim * lag
    base = _draws(rng, 40, dim, nrow, [1.0])
    coef = base.copy()
    target = np.where(np.arange(40) % 2 == 0, 1 - 1e-6, 1 + 1e-6)
    for i in range(40):
        c = _radius(base[i], dim, nrow) / target[i]
        A = base[i, : nrow * dim].reshape(nrow, dim, order="F").copy()
        for l in range(lag):
            A[l * dim:(l + 1) * dim] /= c ** (l + 1)
        coef[i, : nrow * dim] = A.reshape(-1, order="F")
    rho = np.array([_radius(c, dim, nrow) for c in coef])
    np.testing.assert_allclose(rho, target, rtol=1e-9)
    assert np.array_equal(_stable_rows(coef, dim, nrow, None), target < 1)


def test_threshold_and_vhar():
    rng = np.random.default_rng(3)
    dim, week, month = 3, 2, 6
    har = build_vhar(dim, week, month, True)[: 3 * dim, : month * dim]
    coef = _draws(rng, 64, dim, 3 * dim, [0.5, 1.0, 1.6, 2.5])
    rho = np.array([_radius(c, dim, 3 * dim, har) for c in coef])
    assert 5 < (rho < 1).sum() < 59
    assert np.array_equal(_stable_rows(coef, dim, 3 * dim, har), rho < 1)
    assert np.array_equal(_stable_rows(coef, dim, 3 * dim, har, threshold=1.2), rho < 1.2)


def test_etf_sized_vhar_companion():
    """C2 size: 9 series, month = 22 -> 198 x 198 companion matrices."""
    rng = np.random.default_rng(11)
    dim = 9
    har = build_vhar(dim, 5, 22, True)[: 3 * dim, : 22 * dim]
    coef = _draws(rng, 24, dim, 3 * dim, [0.6, 1.0, 1.5])
    rho = np.array([_radius(c, dim, 3 * dim, har) for c in coef])
    assert np.array_equal(_stable_rows(coef, dim, 3 * dim, har), rho < 1)
