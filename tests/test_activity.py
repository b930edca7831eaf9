"""RegRecords::computeActivity (config.h:747-758) as restated by the oracle, against an independent NumPy version of
boost::accumulators::tail_quantile's published rule (left: the ceil(n a)-th smallest sample; right: the
ceil(n (1 - a))-th largest; NaN unless that count is below the number of cached samples), and the shape quirk of the
Select forecasters (forecaster.h:356: the activity VECTOR [alpha | c] is reshaped column-major to d x k)."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle_binding as ob
from bvhar_b200 import _abi
from bvhar_b200._engine import pack_forecast


def numpy_activity(coef_rec, level):
    S = coef_rec.shape[0]
    n_lo = math.ceil(S * (level / 2))
    n_hi = math.ceil(S * (1. - (1 - level / 2)))
    sel = np.empty(coef_rec.shape[1])
    for e in range(coef_rec.shape[1]):
        asc = np.sort(coef_rec[:, e])
        lo = asc[n_lo - 1] if 1 <= n_lo < S else np.nan
        hi = asc[::-1][n_hi - 1] if 1 <= n_hi < S else np.nan
        sel[e] = 0.0 if lo * hi < 0 else 1.1
    return sel


@pytest.mark.parametrize("S,level", [(40, 0.05), (1000, 0.05), (37, 0.5), (8, 0.9)])
def test_oracle_select_forecast_uses_the_published_quantile_rule(S, level):
    rng = np.random.default_rng(S)
    k, p = 2, 2
    nrow = k * p
    alpha = rng.normal(0.0, 1.0, (S, nrow * k)) + rng.normal(0, 1.5, nrow * k)[None, :]
    cst = rng.normal(0.3, 0.2, (S, k))
    rec = {"alpha_record": alpha, "c_record": cst, "a_record": rng.normal(0, .1, (S, 1)), "d_record": np.exp(rng.normal(0, .1, (S, k)))}
    y = rng.normal(size=(10, k))
    D, keep = pack_forecast([rec], [y], 1, p, True, False, [3], level=level)
    out = np.empty(S * k)
    ob.lib().oracle_forecast_density(C.byref(D), 0, ob.dp(out), None)
    D0, keep0 = pack_forecast([rec], [y], 1, p, True, False, [3], level=0.0)
    out0 = np.empty(S * k)
    ob.lib().oracle_forecast_density(C.byref(D0), 0, ob.dp(out0), None)
    # step 1: forecast = mean + noise, the noise is common to both -> the difference is x' ((graph - 1) o coef)
    sel = numpy_activity(np.hstack([alpha, cst]), level)
    d = nrow + 1
    graph = sel.reshape(d, k, order="F")  # forecaster.h:356: column-major reshape of the vector (NOT aligned with coef_mat's rows)
    x = np.concatenate([y[::-1][:p].reshape(-1), [1.0]])
    for i in range(S):
        coef = np.vstack([alpha[i].reshape(nrow, k, order="F"), cst[i][None, :]])
        diff = x @ ((graph - 1.0) * coef)
        got = out.reshape(S, k)[i] - out0.reshape(S, k)[i]
        assert np.allclose(got, diff, rtol=1e-10, atol=1e-12), i
    assert set(np.unique(sel)) <= {0.0, 1.1}
    if S == 1000:  # the two tails use different counts here: ceil(1000 * .025) = 25, ceil(1000 * (1 - .975)) = 26
        assert math.ceil(S * (level / 2)) == 25 and math.ceil(S * (1. - (1 - level / 2))) == 26
