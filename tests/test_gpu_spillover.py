"""bvhar_cuda_spillover (McmcSpilloverRun::returnSpillover, spillover.h:48-81, 246-260; math/structural.h) against the CPU
oracle's literal restatement, draw by draw."""
import numpy as np
import pytest

import oracle_binding as ob
from bvhar_b200._design import build_vhar
from bvhar_b200._engine import spillover

pytestmark = pytest.mark.gpu


def _oracle(coef, contem, svsqrt, dim, nrow, lag, step, har):
    S = coef.shape[0]
    out = {"connect": np.empty((dim, S * dim), order="F"), "net_pairwise": np.empty((dim, S * dim), order="F"),
           "to": np.empty(S * dim), "from": np.empty(S * dim), "tot": np.empty(S)}
    hp = ob.dp(np.asfortranarray(har)) if har is not None else ob.c_dp()
    for i in range(S):
        c = np.empty(dim * dim); n = np.empty(dim * dim); t = np.empty(dim); f = np.empty(dim); tt = np.empty(1)
        a = np.ascontiguousarray(contem[i]) if dim > 1 else np.zeros(1)
        ob.lib().oracle_spillover(dim, nrow, lag, step, ob.dp(np.ascontiguousarray(coef[i])), ob.dp(a), ob.dp(np.ascontiguousarray(svsqrt[i])),
                                  hp, ob.dp(c), ob.dp(t), ob.dp(f), ob.dp(tt), ob.dp(n))
        out["connect"][:, i * dim:(i + 1) * dim] = c.reshape(dim, dim, order="F")
        out["net_pairwise"][:, i * dim:(i + 1) * dim] = n.reshape(dim, dim, order="F")
        out["to"][i * dim:(i + 1) * dim] = t; out["from"][i * dim:(i + 1) * dim] = f; out["tot"][i] = tt[0]
    out["net"] = out["to"] - out["from"]
    return out


def _inputs(rng, S, dim, nrow):
    coef = rng.normal(size=(S, nrow * dim)) * 0.2
    contem = rng.normal(size=(S, max(1, dim * (dim - 1) // 2))) * 0.3
    return coef, contem


@pytest.mark.parametrize("dim,lag,step", [(1, 1, 2), (2, 1, 3), (3, 2, 10), (9, 5, 10), (5, 3, 2), (4, 6, 4), (20, 2, 6)])
def test_var_ldlt_spillover(dim, lag, step):
    rng = np.random.default_rng(dim * 100 + lag * 10 + step)
    S, nrow = 7, dim * lag
    coef, contem = _inputs(rng, S, dim, nrow)
    d = rng.uniform(0.2, 3.0, size=(S, dim))
    got = spillover(coef, contem, d, dim, nrow, lag, step, sv_kind=0)
    ref = _oracle(coef, contem, np.sqrt(d), dim, nrow, lag, step, None)
    for nm in ref:
        np.testing.assert_allclose(got[nm], ref[nm], rtol=1e-10, atol=1e-10, err_msg=nm)
    np.testing.assert_allclose(got["connect"].reshape(dim, dim, S, order="F").sum(axis=1), 100.0, rtol=1e-12)  # rows of a table sum to 100


@pytest.mark.parametrize("kind", [1, 2])
def test_vhar_sv_spillover(kind):
    rng = np.random.default_rng(5 + kind)
    dim, week, month, step, S, n_time = 3, 2, 5, 8, 6, 11
    har = build_vhar(dim, week, month, False)
    coef, contem = _inputs(rng, S, dim, 3 * dim)
    if kind == 1:
        h = rng.normal(size=(S, dim))
        svsqrt = np.exp(h / 2)
    else:
        h = rng.normal(size=(S, n_time * dim))
        svsqrt = np.exp(h.reshape(S, n_time, dim) / 2).mean(axis=1)  # config.h:980-988
    got = spillover(coef, contem, h, dim, 3 * dim, month, step, har_trans=har, sv_kind=kind, n_time=n_time)
    ref = _oracle(coef, contem, svsqrt, dim, 3 * dim, month, step, har)
    for nm in ref:
        np.testing.assert_allclose(got[nm], ref[nm], rtol=1e-10, atol=1e-10, err_msg=nm)


def test_model_spillover_api():
    from bvhar_b200 import datasets
    from bvhar_b200._spec import HorseshoeConfig, LdltConfig, SvConfig
    from bvhar_b200.model import VarBayes, VharBayes
    y = datasets.load_vix()[:120, :4]
    for m in (VarBayes(y, 2, 2, 30, 10, 1, HorseshoeConfig(), LdltConfig(), seed=3), VharBayes(y, 3, 6, 2, 30, 10, 1, HorseshoeConfig(), SvConfig(), seed=3)):
        m.fit()
        sp = m.spillover(5, sparse=False)
        assert set(sp) == {"connect", "net_pairwise", "tot", "to", "from", "net"}
        assert sp["connect"]["mean"].shape == (4, 4) and sp["to"]["mean"].shape == (4,)
        np.testing.assert_allclose(sp["connect"]["mean"].sum(axis=1), 100.0, rtol=1e-9)
        assert np.all(np.isfinite(sp["tot"]["mean"]))
    with pytest.raises(Exception):
        m.spillover(1)


def test_time_sweep_equals_one_call_per_time_point():
    """sv_kind = 3 (DynamicSvSpillover, spillover.h:445-507: one spillover per time point from one set of draws)."""
    rng = np.random.default_rng(9)
    dim, lag, step, S, n_time = 3, 2, 5, 4, 6
    coef, contem = _inputs(rng, S, dim, dim * lag)
    h = rng.normal(size=(S, n_time * dim))
    sweep = spillover(coef, contem, h, dim, dim * lag, lag, step, sv_kind=3, n_time=n_time)
    for t in range(n_time):
        one = spillover(coef, contem, np.ascontiguousarray(h[:, t * dim:(t + 1) * dim]), dim, dim * lag, lag, step, sv_kind=1)
        assert np.array_equal(sweep["tot"][t * S:(t + 1) * S], one["tot"])
        assert np.array_equal(sweep["to"][t * S * dim:(t + 1) * S * dim], one["to"])
        assert np.array_equal(sweep["connect"][:, t * S * dim:(t + 1) * S * dim], one["connect"])


@pytest.mark.parametrize("batch", [None, 1])
@pytest.mark.parametrize("vhar,sv_model", [(False, False), (True, False), (False, True)])
def test_native_dynamic_spillover_matches_fit_plus_spillover(monkeypatch, vhar, sv_model, batch):
    """bvhar_cuda_dynamic_spillover (refit + spillover of every (window, chain), draws kept on the device, windows in
    memory-sized batches) equals fit_run -> records -> bvhar_cuda_spillover on the same window table, bit for bit."""
    import problems as pr
    from bvhar_b200._abi import PackedFit
    from bvhar_b200._engine import CudaFit, dynamic_spillover
    if batch:
        monkeypatch.setenv("BVHAR_MAX_WINDOWS_PER_BATCH", str(batch))
    wins = [(0, 40), (1, 40), (2, 40)]
    kw, meta = pr.make_problem(prior="Horseshoe", sv=sv_model, dim=3, T=60, vhar=vhar, week=2, month=4, chains=2, iters=9, burn=3,
                               thin=2, windows=wins)
    if sv_model:
        kw["lvol_from_init"] = True
    lag = 4 if vhar else 2
    har = build_vhar(3, 2, 4, False) if vhar else None
    to, frm, tot = dynamic_spillover(PackedFit(**kw), 6, lag, har_trans=har)
    assert to.shape == (3, 2, 3, 3) and tot.shape == (3, 2, 3)
    g = CudaFit(PackedFit(**kw)); g.run()
    for w in range(3):
        for c in range(2):
            rec = g.records(w, c)
            if sv_model:
                ref = spillover(rec["alpha_record"], rec["a_record"], rec["h_record"], 3, rec["alpha_record"].shape[1] // 3, lag, 6,
                                har_trans=har, sv_kind=2, n_time=rec["h_record"].shape[1] // 3)
            else:
                ref = spillover(rec["alpha_record"], rec["a_record"], rec["d_record"], 3, rec["alpha_record"].shape[1] // 3, lag, 6, har_trans=har)
            assert np.array_equal(to[w, c].reshape(-1), ref["to"]) and np.array_equal(frm[w, c].reshape(-1), ref["from"])
            assert np.array_equal(tot[w, c], ref["tot"])


def test_model_dynamic_spillover_api():
    from bvhar_b200 import datasets
    from bvhar_b200._spec import HorseshoeConfig, LdltConfig, SvConfig
    from bvhar_b200.model import VarBayes
    y = datasets.load_vix()[:60, :3]
    m = VarBayes(y, 2, 2, 20, 8, 1, HorseshoeConfig(), LdltConfig(), seed=3).fit()
    out = m.dynamic_spillover(55, 4, sparse=False)   # 6 rolling windows
    assert out["tot"]["mean"].shape == (6,) and out["to"]["mean"].shape == (6, 3) and np.all(np.isfinite(out["net"]["upper"]))
    s = VarBayes(y, 2, 2, 20, 8, 1, HorseshoeConfig(), SvConfig(), seed=3).fit()
    out = s.dynamic_spillover(55, 4, sparse=False)   # SV: one spillover per time point of the design
    assert out["tot"]["mean"].shape == (58,) and out["from"]["lower"].shape == (58, 3)
