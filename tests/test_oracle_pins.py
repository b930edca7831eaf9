"""Pins of the CPU oracle and of the host-side data preparation.

The reference's own tests hold no numeric expectation on the sampler ("parity
unpinned"), so the oracle is pinned as far as anything can pin it here:

* the reference's known-answer tables that do exist (``build_grpmat``, python/tests/test_misc.py:6-35);
* the published Random123 known-answer vectors of Philox4x32-10 (the addressed stream that replaces
  boost::random::mt19937, triangular.h:207);
* closed forms checked against NumPy / SciPy (LAPACK): draw_coef (draw.h:372-383), the Minnesota dummy-observation
  moments (design.h:60-98, config.h:144-151) and the MNIW closed form (bayes/mniw/minnesota.h:153-186);
* the identity faithful == efficient (Kronecker design vs weighted Gram, dense vs tridiagonal SV precision).
"""
import ctypes as C

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import oracle_binding as ob
import problems as pr
from bvhar_b200 import _design as dz


# ---------------------------------------------------------------- reference known answers
def test_build_grpmat_reference_tables():
    """python/tests/test_misc.py:6-35 of the reference, verbatim expectations."""
    assert_array_equal(dz.build_grpmat(3, 2, "longrun"), np.array([[2, 1], [1, 2], [4, 3], [3, 4], [6, 5], [5, 6]]))
    assert_array_equal(dz.build_grpmat(3, 2, "short"), np.array([[2, 1], [1, 2], [3, 3], [3, 3], [4, 4], [4, 4]]))
    assert_array_equal(dz.build_grpmat(3, 2, "no"), np.ones((6, 2)))
    with pytest.raises(ValueError):
        dz.build_grpmat(3, 2, "bogus")


def test_philox_known_answer_vectors():
    """Random123 kat_vectors, philox4x32 10 rounds."""
    L = ob.lib()
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, exp in kat:
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        L.oracle_philox4x32_10(c, k, o)
        assert tuple(o) == exp


# ---------------------------------------------------------------- design builders (design.h:8-58)
def test_design_builders():
    rng = np.random.default_rng(0)
    y = rng.normal(size=(30, 3))
    p = 4
    y0 = dz.build_response(y, p, p + 1)
    x0 = dz.build_design(y, p, True)
    assert y0.shape == (26, 3) and x0.shape == (26, 13)
    assert_array_equal(y0, y[p:])
    for lag in range(1, p + 1):  # X0 = [Y_{t-1}, ..., Y_{t-p}, 1]
        assert_array_equal(x0[:, (lag - 1) * 3: lag * 3], y[p - lag: 30 - lag])
    assert_array_equal(x0[:, -1], 1.0)
    T = dz.build_vhar(3, 2, 4, True)
    assert T.shape == (10, 13)
    x1 = dz.build_vhar_design(y, 2, 4, True)
    assert_allclose(x1[:, 0:3], y[3:29])                                   # daily
    assert_allclose(x1[:, 3:6], (y[3:29] + y[2:28]) / 2)                   # weekly mean
    assert_allclose(x1[:, 6:9], (y[3:29] + y[2:28] + y[1:27] + y[0:26]) / 4)  # monthly mean
    assert_array_equal(x1[:, 9], 1.0)


# ---------------------------------------------------------------- draw_coef (draw.h:372-383)
@pytest.mark.parametrize("n,d", [(20, 3), (50, 17), (120, 61)])
def test_draw_coef_matches_lapack(n, d):
    rng = np.random.default_rng(n + d)
    x = np.asfortranarray(rng.normal(size=(n, d))); y = rng.normal(size=n)
    mean = rng.normal(size=d); prec = rng.uniform(0.1, 3.0, d); z = rng.normal(size=d)
    out = np.empty(d)
    ob.lib().oracle_draw_coef(n, d, ob.dp(x), ob.dp(y), ob.dp(mean), ob.dp(prec), ob.dp(z), ob.dp(out))
    P = x.T @ x + np.diag(prec)
    Lc = np.linalg.cholesky(P)
    ref = np.linalg.solve(P, prec * mean + x.T @ y) + np.linalg.solve(Lc.T, z)
    assert_allclose(out, ref, rtol=1e-10, atol=1e-12)
    # the packed-lower entry used by the CUDA building block
    packed = np.concatenate([P[a, : a + 1] for a in range(d)])
    out2 = np.empty(d)
    ob.lib().oracle_draw_coef_packed(d, ob.dp(packed), ob.dp(np.ascontiguousarray(prec * mean + x.T @ y)), ob.dp(z), ob.dp(out2))
    assert_allclose(out2, ref, rtol=1e-10, atol=1e-12)


# ---------------------------------------------------------------- Minnesota moments + MNIW closed form
def test_minnesota_moments_against_dummy_observation_algebra():
    k, lag = 3, 2
    sigma = np.array([0.5, 1.0, 2.0]); lam, eps = 0.2, 1e-4
    daily = np.array([0.9, 0.8, 0.7]); weekly = np.zeros(k); monthly = np.zeros(k)
    nc = k * lag
    mean = np.empty((nc, k), order="F"); full = np.empty((nc, nc), order="F"); aprec = np.empty(k * nc)
    ob.lib().oracle_minnesota_moments(k, lag, ob.dp(sigma), lam, eps, ob.dp(daily), ob.dp(weekly), ob.dp(monthly),
                                      ob.dp(mean), ob.dp(full), ob.dp(aprec))
    # Xd'Xd is diagonal with (l sigma_i / lambda)^2; the prior mean is delta on the first own lag
    expect_diag = np.concatenate([((l + 1) * sigma / lam) ** 2 for l in range(lag)])
    assert_allclose(np.diag(full), expect_diag, rtol=1e-12)
    assert_allclose(full - np.diag(np.diag(full)), 0, atol=1e-14)
    exp_mean = np.zeros((nc, k)); exp_mean[:k, :] = np.diag(daily)
    assert_allclose(mean, exp_mean, atol=1e-12)
    assert_allclose(aprec, np.kron(1 / sigma, expect_diag), rtol=1e-12)
    # the package's own host-side builder (what the engine is fed) agrees with the oracle
    m2, p2 = dz.minnesota_moments({"p": lag, "sigma": sigma, "lambda": lam, "eps": eps, "delta": daily}, k, nc)
    assert_allclose(m2, mean, atol=1e-12)
    assert_allclose(p2, aprec, rtol=1e-12)


def test_mniw_closed_form_posterior():
    """Minnesota::estimateCoef / estimateCov (bayes/mniw/minnesota.h:153-186) on dummy-augmented data
    equals the textbook conjugate update; 1e-10 as north_star asks for deterministic stages."""
    rng = np.random.default_rng(3)
    k, lag, n = 3, 2, 40
    y = rng.normal(size=(n + lag, k)).cumsum(axis=0) * 0.1
    x0 = dz.build_design(y, lag, True); y0 = dz.build_response(y, lag, lag + 1)
    sigma = y.std(axis=0); lam, eps = 0.3, 1e-4; delta = np.full(k, 0.5)
    post = dz.mniw_posterior(x0, y0, lag, sigma, lam, delta, eps)
    xd, yd = post["x_dummy"], post["y_dummy"]
    prior_prec = xd.T @ xd
    prior_mean = np.linalg.solve(prior_prec, xd.T @ yd)
    prec = prior_prec + x0.T @ x0
    coef = np.linalg.solve(prec, prior_prec @ prior_mean + x0.T @ y0)
    assert_allclose(post["prec"], prec, rtol=1e-10)
    assert_allclose(post["coef"], coef, rtol=1e-10, atol=1e-12)
    e0 = yd - xd @ prior_mean
    scale = e0.T @ e0 + y0.T @ y0 + prior_mean.T @ prior_prec @ prior_mean - coef.T @ prec @ coef
    assert_allclose(post["scale"], scale, rtol=1e-8, atol=1e-10)
    assert post["shape"] == xd.shape[0] - x0.shape[1] + 2 + n


# ---------------------------------------------------------------- faithful == efficient
@pytest.mark.parametrize("prior", pr.PRIORS)
@pytest.mark.parametrize("sv", [True, False])
def test_faithful_equals_efficient(prior, sv):
    """The reference's literal algebra (Kronecker design tri.h:286, dense n x n SV precision
    draw.h:454-487) and the weighted-Gram / tridiagonal form the CUDA engine implements give the same
    trajectory under the same addressed random numbers."""
    pk, _ = pr.pack(prior=prior, sv=sv, dim=3, T=36, chains=2, iters=4, burn=1)
    a = ob.OracleFit(pk, faithful=True); b = ob.OracleFit(pk, faithful=False)
    a.run(); b.run()
    assert a.record_names() == b.record_names()
    for nm in a.record_names():
        for c in range(2):
            ra, rb = a.record(nm, 0, c), b.record(nm, 0, c)
            assert ra.shape == rb.shape and ra.shape[0] == 3
            assert_allclose(ra, rb, rtol=2e-8, atol=1e-10, equal_nan=True, err_msg=f"{prior} {nm}")
    a.close(); b.close()


@pytest.mark.parametrize("vhar,mean,ggl", [(True, True, False), (False, False, True), (True, False, True)])
def test_faithful_equals_efficient_variants(vhar, mean, ggl):
    pk, _ = pr.pack(prior="Horseshoe", sv=True, dim=2, T=44, vhar=vhar, week=2, month=4, include_mean=mean, ggl=ggl,
                    chains=1, iters=3, burn=0)
    a = ob.OracleFit(pk, faithful=True); b = ob.OracleFit(pk, faithful=False)
    a.run(); b.run()
    for nm in a.record_names():
        assert_allclose(a.record(nm), b.record(nm), rtol=2e-8, atol=1e-10, equal_nan=True, err_msg=nm)


# ---------------------------------------------------------------- record contract (config.h:581-1251)
EXPECTED_PRIOR_RECORDS = {"Minnesota": [], "HMN": [], "SSVS": ["gamma_record"], "GDP": [],
                          "Horseshoe": ["lambda_record", "eta_record", "tau_record", "kappa_record"],
                          "NG": ["lambda_record", "eta_record", "tau_record"], "DL": ["lambda_record", "tau_record"]}


@pytest.mark.parametrize("prior", pr.PRIORS)
@pytest.mark.parametrize("sv", [True, False])
def test_record_names_and_shapes(prior, sv):
    """Names asserted by the reference's smoke tests (tests/testthat/test-var-bayes.R:31-32, 49, 82, 99)
    and shapes of config.h; kept draws = ceil((iter - burn) / thin) per chain (test-var-bayes.R:251-254)."""
    pk, meta = pr.pack(prior=prior, sv=sv, dim=3, T=30, chains=2, iters=7, burn=2, thin=2)
    f = ob.OracleFit(pk); f.run()
    names = f.record_names()
    base = ["alpha_record", "a_record", "c_record"] + (["h_record", "h0_record", "sigh_record"] if sv else ["d_record"])
    for nm in base + ["alpha_sparse_record", "a_sparse_record", "c_sparse_record"] + EXPECTED_PRIOR_RECORDS[prior]:
        assert nm in names, nm
    k, d, n = 3, pk.dim_design, int(pk.win_nrow[0])
    N, q = k * (d - 1), 3
    kept = 3  # rows 0, 2, 4 of the 5 kept draws (thin_record draw.h:1297-1310)
    assert f.record("alpha_record").shape == (kept, N)
    assert f.record("a_record").shape == (kept, q)
    assert f.record("c_record").shape == (kept, k)
    if sv:
        assert f.record("h_record").shape == (kept, n * k)
    else:
        assert f.record("d_record").shape == (kept, k)
    # thinning keeps draws 0, thin, 2 thin of the unthinned run
    pk1, _ = pr.pack(prior=prior, sv=sv, dim=3, T=30, chains=2, iters=7, burn=2, thin=1)
    g = ob.OracleFit(pk1); g.run()
    assert_allclose(f.record("alpha_record"), g.record("alpha_record")[::2], rtol=0, atol=0)


def test_oracle_sweep_is_reproducible_and_seed_sensitive():
    pk, _ = pr.pack(prior="DL", sv=True, chains=2, iters=4, burn=0)
    a = ob.OracleFit(pk); b = ob.OracleFit(pk)
    a.run(); b.run()
    assert_array_equal(a.record("alpha_record", 0, 1), b.record("alpha_record", 0, 1))
    assert np.abs(a.record("alpha_record", 0, 0) - a.record("alpha_record", 0, 1)).max() > 1e-6


def test_window_table_matches_explicit_slices():
    """Rolling windows as row ranges of one design (forecaster.h:851-859) == fitting each slice by itself."""
    kw, meta = pr.make_problem(prior="NG", sv=True, dim=2, T=40, chains=2, iters=3, burn=0, windows=[(0, 30), (3, 30), (8, 30)])
    from bvhar_b200._abi import PackedFit
    f = ob.OracleFit(PackedFit(**kw)); f.run()
    for w, (r0, nr) in enumerate([(0, 30), (3, 30), (8, 30)]):
        kw1 = dict(kw)
        kw1.pop("win_row0"); kw1.pop("win_nrow")
        kw1["x"] = meta["x"][r0:r0 + nr]; kw1["y"] = meta["resp"][r0:r0 + nr]
        kw1["seed_chain"] = np.asarray(kw["seed_chain"])[w]
        kw1["unit_id_base"] = w * 2
        g = ob.OracleFit(PackedFit(**kw1)); g.run()
        for c in range(2):
            assert_allclose(f.record("alpha_record", w, c), g.record("alpha_record", 0, c), rtol=1e-12, atol=1e-14)
            assert_allclose(f.record("h_record", w, c), g.record("h_record", 0, c), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("dim,lag,step", [(2, 1, 3), (3, 2, 10), (5, 4, 6)])
def test_oracle_spillover_against_companion_form(dim, lag, step):
    """oracle_spillover (literal restatement of spillover.h:48-66 + math/structural.h) against an independent derivation:
    Phi_h = top-left block of the h-th power of the companion matrix, Sigma = L^-1 D L^-T, generalised FEVD
    theta_ij = sigma_jj^-1 sum_h (Phi_h Sigma)_ij^2 / sum_h (Phi_h Sigma Phi_h')_ii (Diebold-Yilmaz), rows normalised."""
    rng = np.random.default_rng(dim + 10 * lag)
    nrow = dim * lag
    A = rng.normal(size=(nrow, dim)) * 0.25
    a = rng.normal(size=max(1, dim * (dim - 1) // 2)) * 0.3
    d = rng.uniform(0.3, 2.0, size=dim)
    L = np.eye(dim)
    for i in range(dim):
        for j in range(i):
            L[i, j] = a[i * (i - 1) // 2 + j]
    Li = np.linalg.inv(L)
    Sigma = Li @ np.diag(d) @ Li.T
    F = np.zeros((nrow, nrow)); F[:dim, :] = A.T; F[dim:, :nrow - dim] = np.eye(nrow - dim)
    num = np.zeros((dim, dim)); den = np.zeros(dim)
    P = np.eye(nrow)
    for h in range(step):
        Phi = P[:dim, :dim]
        PS = Phi @ Sigma
        num += PS ** 2 / np.diag(Sigma)[None, :]
        den += np.diag(PS @ Phi.T)
        P = F @ P
    theta = num / den[:, None]
    sp = 100 * theta / theta.sum(axis=1, keepdims=True)
    off = sp - np.diag(np.diag(sp))
    c = np.empty(dim * dim); n = np.empty(dim * dim); t = np.empty(dim); f = np.empty(dim); tt = np.empty(1)
    ob.lib().oracle_spillover(dim, nrow, lag, step, ob.dp(np.ascontiguousarray(A.reshape(-1, order="F"))), ob.dp(a), ob.dp(np.sqrt(d)),
                              ob.c_dp(), ob.dp(c), ob.dp(t), ob.dp(f), ob.dp(tt), ob.dp(n))
    assert_allclose(c.reshape(dim, dim, order="F"), sp, rtol=1e-10)
    assert_allclose(t, off.sum(axis=0), rtol=1e-10, atol=1e-12)
    assert_allclose(f, off.sum(axis=1), rtol=1e-10, atol=1e-12)
    assert_allclose(tt[0], off.sum() / dim, rtol=1e-10)
    assert_allclose(n.reshape(dim, dim, order="F"), (sp.T - sp) / dim, rtol=1e-9, atol=1e-11)
