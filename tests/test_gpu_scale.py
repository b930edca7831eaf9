"""Size-independent properties at BASELINE.json's full batch sizes (1 024 chains of C3): determinism, invariance
of every unit's draws to how the batch is split (= to the GPU count), stage identities that hold at any size
(Z = Y - XA, L unit lower, SV weights positive), finite posteriors.  The oracle comparison at the same shapes
(a few units of each configuration, stage by stage) is tests/test_gpu_baseline_parity.py."""
import numpy as np
import pytest

import bench
from bvhar_b200._abi import REC_COEF, REC_LVOL_LAST, PackedFit
from bvhar_b200._engine import CudaFit
from bvhar_b200._partition import shard_fit_kwargs, shard_units

pytestmark = pytest.mark.gpu


def _c3(chains, iters, burn, seed=1, prior="DL", record_mask=REC_COEF | REC_LVOL_LAST):
    m = bench.build_model(chains, iters, burn, seed=seed, device=0, prior=prior)
    kw = dict(num_chains=chains, num_iter=iters, num_burn=burn, thin=1, x=m.design_, y=m.response_,
              param_cov=m.cov_spec_.to_dict(), param_prior=m.spec_.to_dict(), param_intercept=m.intercept_spec_.to_dict(),
              param_init=m.init_, prior_type=int(m._prior_type), grp_id=m._group_id, own_id=m._own_id, cross_id=m._cross_id,
              grp_mat=m.group_, include_mean=True, seed_chain=m._seeds(1, chains), is_group=True, record_mask=record_mask)
    return m, kw


def test_c3_full_batch_determinism_and_split_invariance():
    """C3 (k=20, d=61, n=1000), 1,024 chains: two runs are bit-identical, and chains 0..511 / 512..1023 run as two
    separate engines (what 2 GPUs would do) reproduce the joint run bit for bit."""
    m, kw = _c3(1024, 6, 2)
    runs = []
    for _ in range(2):
        f = CudaFit(PackedFit(**kw)); f.run()
        runs.append({nm: f.record_all(nm) for nm in ("alpha_record", "a_record", "h_last_record", "sigh_record")})
        f.close()
    for nm in runs[0]:
        assert np.array_equal(runs[0][nm], runs[1][nm]), nm
        assert np.all(np.isfinite(runs[0][nm])), nm
    parts = []
    for r in range(2):
        sh = shard_units(1, 1024, 2, r)
        f = CudaFit(PackedFit(**shard_fit_kwargs(kw, sh))); f.run()
        parts.append({nm: f.record_all(nm) for nm in runs[0]})
        f.close()
    for nm in runs[0]:
        assert np.array_equal(np.concatenate([parts[0][nm], parts[1][nm]]), runs[0][nm]), nm


def test_c3_stage_identities():
    m, kw = _c3(296, 4, 0)
    f = CudaFit(PackedFit(**kw)); f.step(3, False)
    X, Y = np.asarray(m.design_), np.asarray(m.response_)
    n, k = Y.shape
    for c in (0, 17, 295):
        A = f.get_state("coef_mat", 0, c).reshape(X.shape[1], k, order="F")
        L = f.get_state("chol_lower", 0, c).reshape(k, k, order="F")
        assert np.allclose(np.diag(L), 1.0) and np.allclose(np.triu(L, 1), 0.0)
        a = f.get_state("contem_coef", 0, c)
        assert np.allclose(L[np.tril_indices(k, -1)], a)           # build_inv_lower, draw.h:124-132 (row-wise fill)
    # latent_innov = Y - X A right after updateLatent (stage 6), tri.h:342
    f.run_stage(6, 9)
    for c in (0, 295):
        A = f.get_state("coef_mat", 0, c).reshape(X.shape[1], k, order="F")
        Z = f.get_state("latent_innov", 0, c).reshape(n, k, order="F")
        assert np.max(np.abs(Z - (Y - X @ A))) < 1e-10 * (1 + np.abs(Y).max())
    f.run_stage(3, 9)
    s = f.get_state("sqrt_sv", 0, 5).reshape(n, k, order="F")
    h = f.get_state("lvol_draw", 0, 5).reshape(n, k, order="F")
    assert np.allclose(s, np.exp(h / 2), rtol=1e-12)                # updateSv, tri.h:409


@pytest.mark.parametrize("cfg", ["C1", "C2"])
def test_etf_vix_configs_run_and_mix(cfg):
    """BASELINE configs[0] / configs[1] shapes on the packaged etf_vix data (shortened chains): finite draws,
    sensible posterior (own-lag coefficient of the first series positive), all chains distinct."""
    from bvhar_b200 import _spec as sp
    from bvhar_b200 import datasets
    from bvhar_b200.model import VarBayes, VharBayes
    y = datasets.load_vix()
    if cfg == "C1":
        m = VarBayes(y, 5, 2, 400, 200, 1, sp.SsvsConfig(), sp.LdltConfig(), seed=5)
    else:
        m = VharBayes(y, 5, 22, 8, 400, 200, 1, sp.HorseshoeConfig(), sp.SvConfig(), seed=5)
    m.fit()
    a = m.param_["alpha_record"]
    assert np.all(np.isfinite(a)) and a.shape[1] == 200
    assert m.coef_[0, 0] > 0.3
    assert np.abs(a[0] - a[1]).max() > 1e-8


def test_large_k_cholesky_stress_shape():
    """A cut of C5 (GDP, LDLT, large d): d = 3*30+1 = 91 > 64 exercises three register blocks of the warp
    Cholesky; compared against the oracle on 2 chains."""
    import oracle_binding as ob
    import problems as pr
    pk, _ = pr.pack(prior="GDP", sv=False, dim=30, T=160, vhar=True, week=3, month=6, chains=2, iters=2, burn=0)
    o = ob.OracleFit(pk); g = CudaFit(pk)
    o.step(2, True); g.step(2, True)
    for nm in ("alpha_record", "a_record", "d_record"):
        for c in range(2):
            a, b = o.record(nm, 0, c), g.record(nm, 0, c)
            assert np.max(np.abs(a - b) / (np.abs(a) + 1e-8)) < 1e-7, nm
