"""Conjugate Minnesota samplers on the device (bvhar_cuda_mniw_run, kernels_mniw.cu) against the oracle restatement on the
same addressed Philox stream: fp64, 1e-10 relative (McmcMniw / MhMinnesota, minnesota.h:241-278,
413-510)."""
import numpy as np
import pytest

import oracle_binding as ob
from bvhar_b200 import _abi, minnesota
from mniw_problem import mh_inits, posterior

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-3 * np.abs(b).max() + 1e-300)))


@pytest.mark.parametrize("dim,p,chains,iters,burn,thin", [(3, 2, 2, 12, 4, 1), (9, 5, 3, 20, 5, 3), (20, 3, 2, 8, 0, 1), (1, 1, 2, 6, 1, 2)])
def test_mniw_draws_match_oracle(dim, p, chains, iters, burn, thin):
    _, _, post = posterior(dim=dim, p=p, T=60 + 12 * dim * p // 3)
    seeds = np.arange(chains) + 100
    got = minnesota.estimate_mniw(chains, iters, burn, thin, post["coef"], post["prec"], post["scale"], post["shape"], seeds)
    rc, want = ob.mniw_run(chains, iters, burn, thin, post["coef"], post["prec"], post["scale"], post["shape"], seeds)
    assert rc == 0
    rows = (iters - burn + thin - 1) // thin
    d, k = post["coef"].shape
    for c in range(chains):
        assert got[c]["alpha_record"].shape == (rows, d * k) and got[c]["sigma_record"].shape == (rows, k * k)
        assert _rel(got[c]["sigma_record"], want[c]["sigma_record"]) < 1e-10
        assert _rel(got[c]["alpha_record"], want[c]["alpha_record"]) < 1e-10


def test_mniw_large_design_uses_global_scratch():
    """k = 50, d = 201 (the C4 shape): the d x k work matrix does not fit beside the k x k matrices in shared memory."""
    rng = np.random.default_rng(1)
    k, d = 50, 201
    a = rng.normal(size=(d + 40, d)); prec = a.T @ a
    b = rng.normal(size=(k + 30, k)); scale = b.T @ b
    mean = rng.normal(size=(d, k))
    got = minnesota.estimate_mniw(1, 3, 0, 1, mean, prec, scale, k + 25.0, [7])
    rc, want = ob.mniw_run(1, 3, 0, 1, mean, prec, scale, k + 25.0, [7])
    assert rc == 0
    assert _rel(got[0]["sigma_record"], want[0]["sigma_record"]) < 1e-9
    assert _rel(got[0]["alpha_record"], want[0]["alpha_record"]) < 1e-9


def test_mh_minnesota_matches_oracle():
    x, y0, post = posterior(dim=3, p=2)
    rng = np.random.default_rng(3)
    chains = 3
    par, hess, sv = mh_inits(3, chains, rng)
    prior = {"lambda": {"param": (2.0, 3.0)}, "sigma": {"param": (2.5, 1.2)}}
    inits = [{"par": par[c], "hessian": hess[c], "scale_variance": sv[c]} for c in range(chains)]
    seeds = [21, 22, 23]
    got = minnesota.estimate_bvar_mh(chains, 300, 100, 2, x, y0, post["x_dummy"], post["y_dummy"], prior, inits, seeds)
    G = np.stack([np.linalg.cholesky(sv[c] * np.linalg.inv(hess[c])).T for c in range(chains)])
    rc, want = ob.mniw_run(chains, 300, 100, 2, post["coef"], post["prec"], post["scale"], post["shape"], seeds,
                           mh=dict(par=par, chol_var=G, lambda_param=(2.0, 3.0), psi_param=(2.5, 1.2)))
    assert rc == 0
    for c in range(chains):
        assert list(got[c]) == ["lambda_record", "psi_record", "alpha_record", "sigma_record", "accept_record"]  # minnesota.h:470-478
        assert np.array_equal(got[c]["accept_record"], want[c]["accept_record"])
        assert 0 < got[c]["accept_record"].mean() < 1
        np.testing.assert_allclose(got[c]["lambda_record"], want[c]["lambda_record"], rtol=1e-10)
        np.testing.assert_allclose(got[c]["psi_record"], want[c]["psi_record"], rtol=1e-10)
        assert _rel(got[c]["alpha_record"], want[c]["alpha_record"]) < 1e-10
        assert _rel(got[c]["sigma_record"], want[c]["sigma_record"]) < 1e-10


def test_mniw_errors():
    _, _, post = posterior(dim=3, p=1)
    with pytest.raises(_abi.BvharCudaError, match="shape > dim - 1"):  # sim_iw_tri, random.h:179-181
        minnesota.estimate_mniw(1, 4, 0, 1, post["coef"], post["prec"], post["scale"], 1.5, [1])
    x, y0, _ = posterior(dim=3, p=1)
    prior = {"lambda": {"param": (2.0, 3.0)}, "sigma": {"param": (2.5, 1.2)}}
    inits = [{"par": np.array([0.2, 0.01, 1.0, 1.0]), "hessian": np.eye(4) * 0.2, "scale_variance": 0.05}]
    with pytest.raises(_abi.BvharCudaError, match="should be larger than 0"):  # invgamma_dens, common.h:503-505
        minnesota.estimate_bvar_mh(1, 50, 0, 1, x, y0, post["x_dummy"], post["y_dummy"], prior, inits, [1])


def test_mniw_throughput_smoke():
    """64 chains x 2 000 draws of the C3-sized posterior (k = 20, d = 61): every draw is one CTA."""
    _, _, post = posterior(dim=20, p=3, T=300)
    import time
    t0 = time.perf_counter()
    got = minnesota.estimate_mniw(64, 2000, 1000, 1, post["coef"], post["prec"], post["scale"], post["shape"], np.arange(64) + 1)
    dt = time.perf_counter() - t0
    al = np.stack([g["alpha_record"] for g in got])
    assert al.shape == (64, 1000, 61 * 20) and np.isfinite(al).all()
    print(f"mniw: {64 * 1000 / dt:.0f} kept draws/s end to end")
