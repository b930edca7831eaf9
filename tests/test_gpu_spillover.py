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
