"""The C++ oracle against an independent NumPy restatement of the reference's sweep (tests/numpy_mirror.py).

The mirror is written from the reference's text in its dense formulation (Kronecker design `triangular.h:286`, dense
n x n SV precision `draw.h:454-487`) and reads the UNPACKED constructor arguments, so neither the oracle's algebra nor
the `_abi.PackedFit` packing that feeds both the oracle and the CUDA engine is shared with it.  Only the random variates
are injected (same Philox addresses).  Tolerance 1e-9 relative after each of two full sweeps, every state member."""
import numpy as np
import pytest

import numpy_mirror as nm
import oracle_binding as ob
import problems as pr
from bvhar_b200._abi import PackedFit

COMMON = ["prior_alpha_prec", "coef_mat", "sparse_coef", "prior_chol_prec", "latent_innov", "contem_coef", "sparse_contem", "chol_lower"]
HYPER = {
    "SSVS": ["coef_dummy", "coef_slab", "coef_weight", "spike_scl", "contem_dummy", "contem_slab", "contem_weight", "contem_spike_scl"],
    "DL": ["local_lev", "latent_local", "group_lev", "global_lev", "dir_concen", "contem_local_lev", "latent_contem_local",
           "contem_global_lev", "contem_dir_concen"],
}


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64).reshape(-1); b = np.asarray(b, dtype=np.float64).reshape(-1)
    assert a.shape == b.shape, (a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), "NaN pattern differs"
    ok = ~na
    if not ok.any():
        return 0.0
    scale = np.maximum(np.abs(a[ok]), np.abs(b[ok]))
    floor = 1e-3 * scale.max() + 1e-300
    return float(np.max(np.abs(a[ok] - b[ok]) / np.maximum(scale, floor)))


def mirror_prec_view(m):
    """prior_alpha_prec as the engines expose it: the N slope entries (equation-major) then the k constants."""
    return m.prior_alpha_prec


CASES = [
    ("DL", True, dict(dim=3, T=40, lag=2)),
    ("SSVS", False, dict(dim=3, T=40, lag=2)),
    ("DL", False, dict(dim=2, T=36, lag=1, include_mean=False)),
    ("SSVS", True, dict(dim=2, T=44, vhar=True, week=2, month=4)),
    ("DL", True, dict(dim=4, T=60, lag=3, ggl=False, minnesota=False)),
    ("SSVS", False, dict(dim=5, T=70, lag=2, chains=1)),
]


@pytest.mark.parametrize("prior,sv,kw", CASES, ids=[f"{p}-{'sv' if s else 'ldlt'}-{i}" for i, (p, s, _) in enumerate(CASES)])
@pytest.mark.parametrize("faithful", [True, False], ids=["faithful", "efficient"])
def test_oracle_matches_the_independent_numpy_sweep(prior, sv, kw, faithful):
    kwargs, _ = pr.make_problem(prior=prior, sv=sv, iters=3, burn=0, **kw)
    chains = kwargs["num_chains"]
    o = ob.OracleFit(PackedFit(**kwargs), faithful=faithful)
    mirrors = [nm.MirrorChain(kwargs, c, ob.lib()) for c in range(chains)]
    names = COMMON + HYPER[prior] + (["lvol_draw", "lvol_sig", "lvol_init"] if sv else ["diag_vec"])
    for it in (1, 2):
        o.step(1, False)
        for c, m in enumerate(mirrors):
            m.sweep(it)
            for name in names:
                e = relerr(o.get_state(name, 0, c), m.state(name))
                assert e < 1e-9, f"{prior} sv={sv} faithful={faithful} sweep {it} chain {c} {name}: rel err {e:.3e}"
    o.close()
