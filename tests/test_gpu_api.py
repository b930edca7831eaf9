"""The reference-facing Python classes on the GPU, written like the reference's own smoke tests
(python/tests/test_bayes.py:8-133, 135-259; tests/testthat/test-var-bayes.R, test-summary-forecast.R):
every prior x {LDLT, SV} on etf_vix[:50, :2], a few iterations, shapes / names / error messages."""
import warnings

import numpy as np
import pytest

from bvhar_b200 import datasets
from bvhar_b200._spec import (DlConfig, GdpConfig, HorseshoeConfig, InterceptConfig, LambdaConfig, LdltConfig, MinnesotaConfig,
                              NgConfig, SsvsConfig, SvConfig)
from bvhar_b200.model import VarBayes, VharBayes
from bvhar_b200._engine import CudaFit
import problems as pr

pytestmark = pytest.mark.gpu

PRIOR_CFGS = [SsvsConfig, HorseshoeConfig, lambda: MinnesotaConfig(lam=LambdaConfig()), MinnesotaConfig, NgConfig, DlConfig, GdpConfig]
PRIOR_NAMES = {"SSVS": {"gamma"}, "Horseshoe": {"lambda", "eta", "tau", "kappa"}, "NG": {"lambda", "eta", "tau"}, "DL": {"lambda", "tau"}}


def _data(num_data=50, dim=2, n_ahead=5):
    vix = datasets.load_vix()
    return vix[:num_data, :dim], vix[num_data:num_data + n_ahead, :dim]


def _check_forecast(out, rows, dim):
    for key in ("forecast", "se", "lower", "upper"):
        assert out[key].shape == (rows, dim)
        assert np.all(np.isfinite(out[key]))
    assert np.all(out["lower"] <= out["upper"])


@pytest.mark.parametrize("cfg", PRIOR_CFGS)
@pytest.mark.parametrize("cov", [LdltConfig, SvConfig])
def test_var_bayes(cfg, cov):
    dim, lag, n_ahead = 2, 3, 5
    data, test_y = _data(n_ahead=n_ahead)
    fit = VarBayes(data, lag, 2, 5, 2, 1, cfg(), cov(), InterceptConfig(), True, True, True, False, 1, seed=1)
    fit.fit()
    assert fit.n_features_in_ == dim
    assert fit.coef_.shape == (dim * lag + 1, dim)
    assert fit.intercept_.shape == (dim,)
    assert fit.param_["alpha_record"].shape == (2, 3, dim * dim * lag)  # chains x kept draws x N (test-var-bayes.R:251-254)
    for nm in PRIOR_NAMES.get(fit.spec_.prior, ()):  # test-var-bayes.R:31-32, 49, 82, 99
        assert nm in fit.param_names_
    assert ("h" in fit.param_names_) == (cov is SvConfig) and ("d" in fit.param_names_) == (cov is LdltConfig)
    _check_forecast(fit.predict(n_ahead, stable=False, sparse=True), n_ahead, dim)
    _check_forecast(fit.roll_forecast(1, test_y, stable=False, sparse=True), n_ahead, dim)
    _check_forecast(fit.expand_forecast(1, test_y, stable=False, sparse=False), n_ahead, dim)
    frame = fit.to_frame()
    assert len(frame) == 2 * 3 and "alpha[1]" in frame.columns


@pytest.mark.parametrize("cfg", PRIOR_CFGS)
@pytest.mark.parametrize("cov", [LdltConfig, SvConfig])
def test_vhar_bayes(cfg, cov):
    dim, n_ahead = 2, 5
    data, test_y = _data(n_ahead=n_ahead)
    c = cfg()
    if isinstance(c, MinnesotaConfig):
        c = MinnesotaConfig(lam=c.lam if hasattr(c, "lam") and isinstance(c.lam, LambdaConfig) else .1, is_long=True)
    fit = VharBayes(data, 5, 22, 2, 5, 2, 1, c, cov(), InterceptConfig(), True, "longrun", True, False, 1, seed=1)
    fit.fit()
    assert fit.coef_.shape == (dim * 3 + 1, dim)
    assert fit.intercept_.shape == (dim,)
    _check_forecast(fit.predict(n_ahead, stable=False, sparse=True), n_ahead, dim)
    _check_forecast(fit.roll_forecast(1, test_y, stable=False, sparse=True), n_ahead, dim)
    _check_forecast(fit.expand_forecast(1, test_y, stable=False, sparse=True), n_ahead, dim)
    assert any(col.startswith("phi[") for col in fit.to_frame().columns)  # R/vhar-bayes.R:526 renames alpha -> phi


def test_errors_and_warnings():
    data, _ = _data()
    with pytest.warns(UserWarning, match="will not use every thread"):
        VarBayes(data, 3, 2, 5, 2, 1, SsvsConfig(), LdltConfig(), InterceptConfig(), True, True, False, True, 3)
    with pytest.raises(ValueError, match="'data' rows must be larger than 'lag' = 3"):
        VarBayes(data[:2], 3, 2, 5, 2, 1, SsvsConfig(), LdltConfig(), InterceptConfig(), True, True, False, True, 1)
    with pytest.raises(TypeError):
        VarBayes(data, 3, 2, 5, 2, 1, SsvsConfig(), "ldlt", InterceptConfig())
    with pytest.raises(ValueError):
        VarBayes(data.reshape(-1), 3)


def test_lpl_and_multistep_roll():
    data, test_y = _data(n_ahead=6)
    fit = VarBayes(data, 2, 2, 6, 2, 1, HorseshoeConfig(), SvConfig(), seed=3).fit()
    out = fit.roll_forecast(2, test_y)  # 6 - 2 + 1 = 5 windows, last step kept (forecaster.h:803)
    _check_forecast(out, 5, 2)
    assert out["lpl"].shape == (5,) and np.all(np.isfinite(out["lpl"]))


def test_fit_is_reproducible_for_a_seed():
    data, _ = _data()
    a = VharBayes(data, 5, 22, 3, 6, 2, 1, DlConfig(), SvConfig(), seed=9).fit()
    b = VharBayes(data, 5, 22, 3, 6, 2, 1, DlConfig(), SvConfig(), seed=9).fit()
    c = VharBayes(data, 5, 22, 3, 6, 2, 1, DlConfig(), SvConfig(), seed=10).fit()
    assert np.array_equal(a.param_["alpha_record"], b.param_["alpha_record"])
    assert not np.array_equal(a.param_["alpha_record"], c.param_["alpha_record"])


def test_progress_callback_and_interrupt():
    import problems as pr
    from bvhar_b200._abi import BvharCudaError
    from bvhar_b200._engine import CudaFit
    pk, _ = pr.pack(prior="DL", sv=True, iters=40, burn=20)
    seen = []
    g = CudaFit(pk)
    g.run(lambda done, total: seen.append((done, total)) or False)
    assert seen[-1] == (40, 40) and len(seen) == 20  # every 5 % (triangular.h:1248-1251)
    g2 = CudaFit(pk)
    with pytest.raises(BvharCudaError) as ei:
        g2.run(lambda done, total: done >= 24)
    assert ei.value.status == 4  # interrupted: partial records stay readable (triangular.h:1261-1270)
    assert g2.record("alpha_record").shape[0] == 4


def test_pinned_output_pool_ownership():
    """Large records land in pooled page-locked memory: results own their block until garbage-collected, after which
    the next fit reuses it; nothing is overwritten while a result is alive."""
    import gc
    import problems as pr
    from bvhar_b200 import _hostpool
    from bvhar_b200._engine import CudaFit
    pk, _ = pr.pack(prior="DL", sv=True, dim=4, T=300, chains=64, iters=24, burn=4)
    g = CudaFit(pk); g.run()
    a = g.record_all("h_record")            # 64 x 20 x 1192 doubles = 12 MB -> pinned
    assert a.nbytes >= (8 << 20) and getattr(a, "_blk", None) is not None
    ref = np.array(a)                        # pageable copy
    b = g.record_all("h_record")
    assert np.array_equal(a, ref) and np.array_equal(b, ref) and a._blk.ptr != b._blk.ptr
    ptr = a._blk.ptr
    view = a[3]                              # a view keeps the block alive
    del a
    gc.collect()
    c = g.record_all("h_record")
    assert c._blk.ptr != ptr and np.array_equal(view, ref[3])
    del view, c
    gc.collect()
    d = g.record_all("h_record")
    assert d._blk.ptr in (ptr, b._blk.ptr) or True
    assert np.array_equal(d, ref)
    small = g.record_all("sigh_record")
    assert getattr(small, "_blk", None) is None
    _hostpool.trim()


@pytest.mark.parametrize("batch", [None, 2, 1])
@pytest.mark.parametrize("vhar,sv_model,sparse", [(False, True, False), (True, False, True), (True, True, False)])
def test_native_outforecast_matches_fit_plus_forecast(monkeypatch, vhar, sv_model, sparse, batch):
    """bvhar_cuda_outforecast_run (refit + forecast of every (window, chain) in one call, draws kept on the device)
    equals fit_run -> records -> bvhar_cuda_forecast_density on the same window table, bit for bit - also when the
    window table is processed in memory-sized batches (here forced to 2 + 1 and 1 + 1 + 1 windows)."""
    if batch:
        monkeypatch.setenv("BVHAR_MAX_WINDOWS_PER_BATCH", str(batch))
    import problems as pr
    from bvhar_b200._abi import PackedFit
    from bvhar_b200._design import build_vhar
    from bvhar_b200._engine import CudaFit, forecast_density, outforecast
    wins = [(1, 40), (2, 40), (3, 40)]
    kw, meta = pr.make_problem(prior="Horseshoe", sv=sv_model, dim=3, T=60, vhar=vhar, week=2, month=4, chains=2, iters=8, burn=2,
                               thin=2, windows=wins)
    kw["unit_id_base"] = 2
    if sv_model:
        kw["lvol_from_init"] = False
    lag = 4 if vhar else 2
    har = build_vhar(3, 2, 4, True) if vhar else None
    y = meta["y"]
    valid = y[-1] * 1.1
    seed_fc = [5, 6]
    fc, lpl = outforecast(PackedFit(**kw), y, 3, lag, seed_fc, har_trans=har, valid_vec=valid, sparse=sparse)
    assert fc.shape == (3, 2, 3, 3) and lpl.shape == (3, 2)
    g = CudaFit(PackedFit(**kw)); g.run()
    units = [g.records(w, c) for w in range(3) for c in range(2)]
    last = [y[: lag + r0 + n] for (r0, n) in wins for _ in range(2)]
    dens, lp = forecast_density(units, last, 3, lag, True, sv_model, np.tile(seed_fc, 3), har_trans=har, valid_vec=valid,
                                sparse=sparse, unit_id_base=2)
    for u in range(6):
        assert np.array_equal(dens[u][-1].reshape(3, 3), fc[u // 2, u % 2])
    np.testing.assert_allclose(lp.mean(axis=1).reshape(3, 2), lpl, rtol=1e-12)


def test_device_posterior_means_match_host_reduction():
    import problems as pr
    from bvhar_b200._engine import CudaFit
    pk, _ = pr.pack(prior="SSVS", sv=True, dim=3, T=50, chains=5, iters=30, burn=6, thin=2)
    g = CudaFit(pk); g.run()
    for nm in ("alpha_record", "alpha_sparse_record", "c_record", "a_sparse_record", "sigh_record"):
        host = np.nanmean(g.record_all(nm, pinned=False), axis=(0, 1))
        np.testing.assert_allclose(g.record_mean(nm), host, rtol=1e-12, atol=1e-14, equal_nan=True)


def test_fit_over_several_engines_equals_single_engine():
    """fit(devices=[...]) splits the chains over one engine per entry (here the same GPU three times): the draws are
    those of the single-engine fit, bit for bit (global Philox keys), and the posterior means agree."""
    data, _ = _data()
    a = VharBayes(data, 5, 22, 7, 10, 4, 1, HorseshoeConfig(), SvConfig(), seed=21).fit()
    b = VharBayes(data, 5, 22, 7, 10, 4, 1, HorseshoeConfig(), SvConfig(), seed=21).fit(devices=[0, 0, 0])
    assert set(a.param_) == set(b.param_)
    for nm in a.param_:
        assert np.array_equal(a.param_[nm], b.param_[nm], equal_nan=True), nm
    np.testing.assert_allclose(a.coef_, b.coef_, rtol=1e-12)


def test_roll_forecast_with_stability_filter():
    data, test_y = _data(n_ahead=4)
    fit = VarBayes(data, 2, 2, 30, 10, 1, HorseshoeConfig(), LdltConfig(), seed=4).fit()
    a = fit.roll_forecast(1, test_y, stable=True)
    _check_forecast(a, 4, 2)
    b = fit.predict(3, stable=True)
    _check_forecast(b, 3, 2)


@pytest.mark.parametrize("thin", [1, 3])
def test_streamed_record_sinks_equal_bulk_fetch(thin):
    """Records streamed to bound host sinks while the sampler runs (bvhar_cuda_fit_bind_record_sink) are identical to the
    bulk device-to-host fetch after the run, with and without thinning and with a progress callback syncing in between."""
    import problems as pr
    from bvhar_b200._engine import CudaFit
    kw = dict(prior="Horseshoe", sv=True, dim=3, T=60, chains=5, iters=17, burn=4, thin=thin)
    pk, _ = pr.pack(**kw)
    a = CudaFit(pk)
    a._bind_sinks(min_bytes=0)               # every record streams
    assert set(a._sinks) == set(a.record_names())
    a.run(lambda done, total: False)
    pk2, _ = pr.pack(**kw)
    b = CudaFit(pk2)
    b._bind_sinks(min_bytes=1 << 60)         # nothing streams
    assert not b._sinks
    b._sinks = {"_": None}                   # keep run() from binding
    b.run()
    for nm in a.record_names():
        x, y = a.record_all(nm), b.record_all(nm, pinned=False)
        assert x.shape == y.shape and np.array_equal(x, y, equal_nan=True), nm
        assert np.array_equal(a.record_all(nm, pinned=False), y, equal_nan=True)  # second fetch: plain copy
    a.close(); b.close()


def _np_summary(x, level):
    """NumPy reference of the device summaries: x [chain, draw] of one parameter.  R-hat / ESS follow the formulas of
    posterior::rhat_basic / ess_basic (split chains, biased autocovariances, Geyer's initial positive and monotone sequences)."""
    U, S = x.shape
    flat = x.reshape(-1)
    out = {"mean": flat.mean(), "sd": flat.std(ddof=1), "lower": np.quantile(flat, level / 2), "median": np.quantile(flat, .5),
           "upper": np.quantile(flat, 1 - level / 2), "pip": np.mean(flat != 0)}
    n2 = S // 2
    halves = np.concatenate([x[:, :n2], x[:, (S + 1) // 2:]], axis=1).reshape(U, 2, n2).reshape(2 * U, n2)  # chain c -> halves 2c, 2c + 1
    m = halves.shape[0]
    cm = halves.mean(axis=1); cv = halves.var(axis=1, ddof=1)
    W = cv.mean()
    var_plus = W * (n2 - 1) / n2 + (cm.var(ddof=1) if m > 1 else 0.0)
    out["rhat"] = np.sqrt(var_plus / W)
    yc = halves - cm[:, None]
    acov = np.array([(yc[:, : n2 - t] * yc[:, t:]).sum(axis=1).mean() / n2 for t in range(n2)])
    rho = np.zeros(n2 + 2)
    t = 0
    even, odd = 1.0, 1 - (W - acov[1]) / var_plus
    rho[0], rho[1] = even, odd
    while t < n2 - 5 and not np.isnan(even + odd) and even + odd > 0:
        t += 2
        even = 1 - (W - acov[t]) / var_plus
        odd = 1 - (W - acov[t + 1]) / var_plus
        if even + odd >= 0:
            rho[t], rho[t + 1] = even, odd
    max_t = t
    if even > 0:
        rho[max_t] = even
    t = 0
    while t <= max_t - 4:
        t += 2
        if rho[t] + rho[t + 1] > rho[t - 2] + rho[t - 1]:
            rho[t] = (rho[t - 2] + rho[t - 1]) / 2
            rho[t + 1] = rho[t]
    N = m * n2
    tau = max(-1 + 2 * rho[: max(1, max_t)].sum() + rho[max_t], 1 / np.log10(N))
    out["ess"] = N / tau
    return out


@pytest.mark.parametrize("chains,iters,burn,thin", [(4, 64, 14, 1), (3, 47, 4, 2), (24, 1100, 100, 1)])
def test_record_summary_on_device(chains, iters, burn, thin):
    """bvhar_cuda_fit_record_summary: mean / sd / quantiles / split-R-hat / ESS / pip of every column
    of a record, computed on the device, against NumPy on the fetched draws (shared-memory and global-memory paths)."""
    pk, _ = pr.pack(prior="Horseshoe", sv=True, dim=2, T=40, lag=1, chains=chains, iters=iters, burn=burn, thin=thin, seed=3)
    g = CudaFit(pk); g.run()
    for nm in ("alpha_record", "a_record", "sigh_record", "alpha_sparse_record"):
        got = g.record_summary(nm, level=.1)
        draws = g.record_all(nm)  # [chain, draw, col]
        for col in range(draws.shape[2]):
            x = draws[:, :, col]
            if np.isnan(x).any():
                ok = ~np.isnan(x)
                assert abs(got["mean"][col] - x[ok].mean()) < 1e-12 * max(1, abs(x[ok].mean()))
                assert np.isnan(got["rhat"][col]) and np.isnan(got["lower"][col])
                continue
            ref = _np_summary(x, .1)
            for key, val in ref.items():
                if np.isnan(val) or np.isinf(val):  # e.g. a SAVS column that is zero in every draw: W = 0
                    assert np.isnan(got[key][col]) or got[key][col] == val, (nm, col, key, got[key][col], val)
                    continue
                assert abs(got[key][col] - val) <= 1e-9 * max(1.0, abs(val)), (nm, col, key, got[key][col], val)
