"""The oracle's restatement of the conjugate Minnesota samplers (McmcMniw / MhMinnesota, minnesota.h:241-278, 413-510):
pinned (a) against an independent NumPy / LAPACK evaluation of sim_mn_iw fed the very same addressed variates, and
(b) against the closed-form moments of the matrix-normal inverse-Wishart law."""
import numpy as np
import scipy.linalg as sl

import oracle_binding as ob
from mniw_problem import mh_inits, posterior

ST_CHI, ST_BART, ST_Z, ST_CAND, ST_U = 32, 33, 34, 35, 36


def _numpy_draw(mean, prec, scale, shape, seed, unit, t):
    """sim_mn_iw (math/random.h:177-211) with LAPACK, the variates fetched by address from the oracle's stream."""
    d, k = mean.shape
    L = ob.lib()
    L.oracle_draw_at.restype = __import__("ctypes").c_double
    Q = np.zeros((k, k))
    for i in range(k):
        Q[i, i] = np.sqrt(L.oracle_draw_at(2, seed, unit, t, ST_CHI, i, 0.5 * (shape - i), 2.0, 0.0))
        for j in range(i + 1, k):
            Q[i, j] = L.oracle_draw_at(0, seed, unit, t, ST_BART, i * k + j, 0, 0, 0)
    chol_scale = np.linalg.cholesky(scale)
    res = chol_scale @ sl.solve_triangular(Q.T, np.eye(k), lower=True)
    Sig = res @ res.T
    Z = np.array([[L.oracle_draw_at(0, seed, unit, t, ST_Z, i * k + j, 0, 0, 0) for j in range(k)] for i in range(d)])
    Uv = np.linalg.cholesky(Sig).T
    Uu = np.linalg.cholesky(prec).T
    return mean + sl.solve_triangular(Uu, Z @ Uv, lower=False), Sig


def test_oracle_mniw_draw_matches_lapack():
    _, _, post = posterior(dim=3, p=2)
    seeds = [11, 12]
    rc, rec = ob.mniw_run(2, 6, 2, 2, post["coef"], post["prec"], post["scale"], post["shape"], seeds)
    assert rc == 0
    d, k = post["coef"].shape
    for c in range(2):
        assert rec[c]["alpha_record"].shape == (2, d * k) and rec[c]["sigma_record"].shape == (2, k * k)
        for r, t in enumerate((3, 5)):  # mcmc_step = num_burn + 1 + r * thin
            A, Sig = _numpy_draw(post["coef"], post["prec"], post["scale"], post["shape"], seeds[c], c, t)
            np.testing.assert_allclose(rec[c]["alpha_record"][r], A.reshape(-1, order="F"), rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(rec[c]["sigma_record"][r], Sig.reshape(-1, order="F"), rtol=1e-10, atol=1e-14)


def test_oracle_mniw_moments():
    """Law of the draws against an independent NumPy sampler of the reference's literal construction (Generator chi-square /
    normal variates, LAPACK algebra): E Sigma, E A = mean, Var A_ij = E Sigma_jj (prec^-1)_ii, within 4.5 Monte Carlo s.e.
    NB the reference builds Sigma = L (Q Q')^-1 L' from an UPPER Bartlett factor whose chi-square degrees of freedom
    decrease down the diagonal (random.h:186-199), which is not the textbook inverse Wishart (its mean differs from
    scale / (shape - k - 1) by a few per cent): the restatement keeps that construction, and so does this check."""
    _, _, post = posterior(dim=3, p=1)
    n = 6000
    rc, rec = ob.mniw_run(1, n, 0, 1, post["coef"], post["prec"], post["scale"], post["shape"], [2024])
    assert rc == 0
    d, k = post["coef"].shape
    sig = rec[0]["sigma_record"]; al = rec[0]["alpha_record"]
    g = np.random.default_rng(99)
    Ls = np.linalg.cholesky(post["scale"])
    ref = np.empty((n, k * k))
    for s_ in range(n):
        Q = np.triu(g.normal(size=(k, k)), 1) + np.diag(np.sqrt(g.chisquare(post["shape"] - np.arange(k))))
        res = Ls @ sl.solve_triangular(Q.T, np.eye(k), lower=True)
        ref[s_] = (res @ res.T).reshape(-1, order="F")
    se = np.sqrt(sig.var(axis=0, ddof=1) / n + ref.var(axis=0, ddof=1) / n)
    assert np.all(np.abs(sig.mean(axis=0) - ref.mean(axis=0)) < 4.5 * se)
    assert np.all(np.abs(sig.std(axis=0, ddof=1) / ref.std(axis=0, ddof=1) - 1.0) < 0.1)
    se_a = al.std(axis=0, ddof=1) / np.sqrt(n)
    assert np.all(np.abs(al.mean(axis=0) - post["coef"].reshape(-1, order="F")) < 4.5 * se_a)
    esig = sig.mean(axis=0).reshape(k, k)
    var_th = np.kron(np.diag(esig), np.diag(np.linalg.inv(post["prec"])))
    assert np.all(np.abs(al.var(axis=0, ddof=1) / var_th - 1.0) < 0.12)


def test_oracle_mh_minnesota_records():
    """MhMinnesota: the Metropolis state only moves on acceptance, proposals are centred on the initial point (the reference
    never updates prevprior), and the acceptance rule is log u < min(ratio, 0) of the Gamma / inverse-Gamma hyper-priors."""
    _, _, post = posterior(dim=3, p=1)
    rng = np.random.default_rng(3)
    par, hess, sv = mh_inits(3, 2, rng)
    G = np.stack([np.linalg.cholesky(sv[c] * np.linalg.inv(hess[c])).T for c in range(2)])
    mh = dict(par=par, chol_var=G, lambda_param=(2.0, 3.0), psi_param=(2.5, 1.2))
    rc, rec = ob.mniw_run(2, 400, 0, 1, post["coef"], post["prec"], post["scale"], post["shape"], [5, 6], mh=mh)
    assert rc == 0
    import ctypes
    L = ob.lib()
    from scipy.stats import gamma as gdist, invgamma
    for c in range(2):
        lam, psi, acc = rec[c]["lambda_record"], rec[c]["psi_record"], rec[c]["accept_record"]
        assert 0.05 < acc.mean() < 1.0
        prev = par[c]
        state = prev.copy()
        for r in range(400):
            t = r + 1
            z = np.array([L.oracle_draw_at(0, [5, 6][c], c, t, ST_CAND, i, 0, 0, 0) for i in range(4)])
            cand = prev + z @ G[c]
            # gamma_shp <- rate, gamma_rate <- shape (minnesota.h:426): Gamma(shape = 3, scale = 1 / 2)
            lr = (gdist.logpdf(cand[0], 3.0, scale=0.5) + invgamma.logpdf(cand[1:], 2.5, scale=1.2).sum()
                  - gdist.logpdf(prev[0], 3.0, scale=0.5) - invgamma.logpdf(prev[1:], 2.5, scale=1.2).sum())
            u = L.oracle_draw_at(6, [5, 6][c], c, t, ST_U, 0, 0, 0, 0)
            want = np.log(u) < min(lr, 0.0)
            assert bool(acc[r]) == bool(want)
            if want:
                state = cand
            np.testing.assert_allclose(np.r_[lam[r], psi[r]], state, rtol=1e-12)


def test_oracle_mh_negative_psi_is_an_error():
    """A proposal with a negative psi makes invgamma_dens STOP in the reference (common.h:503-505): status 3."""
    _, _, post = posterior(dim=3, p=1)
    par = np.array([[0.2, 0.01, 1.0, 1.0]])
    G = np.eye(4)[None] * 0.5
    rc, _ = ob.mniw_run(1, 50, 0, 1, post["coef"], post["prec"], post["scale"], post["shape"], [1],
                        mh=dict(par=par, chol_var=G, lambda_param=(2.0, 3.0), psi_param=(2.5, 1.2)))
    assert rc == 3
