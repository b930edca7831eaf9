"""Host-side mirror of the reference's sampler classes, backed by the CUDA engine's C-ABI.

Classes and constructor signatures follow the pybind11 bindings of the reference
(python/src/bvhar/_src/_cta.cpp:4-38 ``McmcLdlt``, ``McmcLdltGrp``, ``SvMcmc``, ``SvGrpMcmc``;
:40-208 ``LdltForecast`` ... ``SvGrpVharExpand``), i.e. the template instantiations of
``bvhar::McmcRun<BaseMcmc, isGroup>`` (triangular.h:1192-1295), ``McmcForecastRun`` and
``Mcmc{Var,Vhar}forecastRun`` (forecaster.h:494-1121).  ``returnRecords()`` returns one dict per chain
with the record names of config.h:639-648, 865-867, 967-971, 814-820.

There is no CPU fallback: constructing any of these without the built CUDA library or without a
GPU raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _abi
from ._abi import BvharCudaError, ForecastDesc, PackedFit, REC_ALL, c_dp, c_up, check, load_library
from ._design import build_design, build_response, build_vhar


class CudaFit:
    """A batch of (window, chain) sampling units living on one GPU."""

    def __init__(self, packed: PackedFit):
        self.packed = packed
        self.L = load_library()
        self._h = C.c_void_p()
        check(self.L.bvhar_cuda_fit_create(C.byref(packed.desc), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.L.bvhar_cuda_fit_destroy(self._h)
            self._h = C.c_void_p()
        self._sink_keep = []  # the engine no longer holds the sink pointers
        self._sinks = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- driving ----------------------------------------------------------------------------
    def run(self, progress=None):
        """num_burn warm-up sweeps, then num_iter - num_burn recorded sweeps (triangular.h:1241-1284)."""
        if progress is None:
            cb = _abi.PROGRESS_FN()
        else:
            cb = _abi.PROGRESS_FN(lambda ctx, done, total: int(bool(progress(done, total))))
        self._bind_sinks()
        check(self.L.bvhar_cuda_fit_run(self._h, cb, None))

    def _bind_sinks(self, min_bytes: int = 8 << 20):
        """Large records stream to pooled page-locked host buffers while the sampler runs (bvhar_cuda_fit_bind_record_sink);
        record_all() then hands the buffer back without a device-to-host pass at the end."""
        pk = self.packed
        thin = max(1, int(pk.desc.thin))
        kept = (max(0, int(pk.desc.num_iter) - int(pk.desc.num_burn)) + thin - 1) // thin
        U = pk.n_windows * pk.n_chains
        if kept < 1 or getattr(self, "_sinks", None):
            return
        self._sinks = {}
        from . import _hostpool
        for nm in self.record_names():
            r, c = C.c_int64(), C.c_int64()
            check(self.L.bvhar_cuda_fit_record_dims(self._h, nm.encode(), C.byref(r), C.byref(c)))
            if r.value != 0 or U * kept * c.value * 8 < min_bytes:
                continue  # (already recording, or small: fetched at the end)
            out = _hostpool.empty((U, kept, c.value))
            check(self.L.bvhar_cuda_fit_bind_record_sink(self._h, nm.encode(), out.ctypes.data_as(c_dp), U, kept, c.value))
            self._sinks[nm] = out

    def step(self, n: int = 1, record: bool = False, sync: bool = True):
        check(self.L.bvhar_cuda_fit_step(self._h, int(n), int(record)))
        if sync:
            self.sync()

    def sync(self):
        check(self.L.bvhar_cuda_fit_sync(self._h))

    def run_stage(self, stage: int, it: int):
        check(self.L.bvhar_cuda_fit_run_stage(self._h, int(stage), int(it)))

    def last_ms(self) -> float:
        v = C.c_double()
        check(self.L.bvhar_cuda_fit_last_ms(self._h, C.byref(v)))
        return v.value

    def set_profiling(self, on: bool):
        check(self.L.bvhar_cuda_fit_set_profiling(self._h, int(on)))

    def profile(self) -> Dict[str, Any]:
        """{kernel name: (launches, total ms)} accumulated while profiling was on."""
        buf = C.create_string_buffer(8192)
        check(self.L.bvhar_cuda_fit_profile(self._h, buf, 8192))
        out = {}
        for line in buf.value.decode().splitlines():
            nm, cnt, ms = line.split()
            out[nm] = (int(cnt), float(ms))
        return out

    def launch_count(self) -> int:
        v = C.c_int64()
        check(self.L.bvhar_cuda_fit_launch_count(self._h, C.byref(v)))
        return v.value

    # -- state (parity tests) ---------------------------------------------------------------
    def get_state(self, name: str, window: int = 0, chain: int = 0) -> np.ndarray:
        n = C.c_int64()
        check(self.L.bvhar_cuda_fit_state_len(self._h, name.encode(), C.byref(n)))
        ln = n.value
        if name in ("latent_innov", "lvol_draw", "sqrt_sv"):
            ln = int(self.packed.win_nrow[window]) * self.packed.dim
        out = np.empty(ln)
        check(self.L.bvhar_cuda_fit_get_state(self._h, window, chain, name.encode(), out.ctypes.data_as(c_dp), ln))
        return out

    def set_state(self, name: str, value, window: int = 0, chain: int = 0):
        v = np.ascontiguousarray(np.asarray(value, dtype=np.float64).reshape(-1, order="F"))
        check(self.L.bvhar_cuda_fit_set_state(self._h, window, chain, name.encode(), v.ctypes.data_as(c_dp), v.size))

    # -- records ----------------------------------------------------------------------------
    def record_names(self) -> List[str]:
        buf = C.create_string_buffer(4096)
        check(self.L.bvhar_cuda_fit_record_names(self._h, buf, 4096))
        return [s for s in buf.value.decode().split("\n") if s]

    def record(self, name: str, window: int = 0, chain: int = 0) -> np.ndarray:
        r, c = C.c_int64(), C.c_int64()
        check(self.L.bvhar_cuda_fit_record_dims(self._h, name.encode(), C.byref(r), C.byref(c)))
        out = np.empty((r.value, c.value), order="F")
        check(self.L.bvhar_cuda_fit_record(self._h, window, chain, name.encode(), out.ctypes.data_as(c_dp), r.value, c.value))
        return out

    def record_all(self, name: str, pinned: bool = True) -> np.ndarray:
        """ndarray[unit, draw, column] of one record for every unit (one bulk device-to-host copy, into pooled
        page-locked memory for large records)."""
        r, c = C.c_int64(), C.c_int64()
        check(self.L.bvhar_cuda_fit_record_dims(self._h, name.encode(), C.byref(r), C.byref(c)))
        U = self.packed.n_windows * self.packed.n_chains
        shape = (U, r.value, c.value)
        sinks = getattr(self, "_sinks", None) or {}
        sink = sinks.get(name)
        if sink is not None and sink.shape == shape:  # streamed during run(): only waits for the tail of the copy stream
            check(self.L.bvhar_cuda_fit_record_all(self._h, name.encode(), sink.ctypes.data_as(c_dp), U, r.value, c.value))
            # handed out once; the engine keeps the pointer bound until it is destroyed, so the block must not go back
            # to the pool before close(): a reference stays in _sink_keep
            self._sink_keep = getattr(self, "_sink_keep", [])
            self._sink_keep.append(sinks.pop(name))
            return sink
        # (a sink whose shape does not match - a run cut short by the progress callback keeps fewer rows - stays bound
        # and alive: a resumed run() still streams into it; the rows kept so far are fetched from the device below)
        if pinned and U * r.value * c.value * 8 >= (8 << 20):
            from . import _hostpool
            out = _hostpool.empty(shape)
        else:
            out = np.empty(shape)
        check(self.L.bvhar_cuda_fit_record_all(self._h, name.encode(), out.ctypes.data_as(c_dp), U, r.value, c.value))
        return out

    def record_mean(self, name: str) -> np.ndarray:
        """Posterior mean of each column over all units and kept draws (NaNs skipped), reduced on the device."""
        r, c = C.c_int64(), C.c_int64()
        check(self.L.bvhar_cuda_fit_record_dims(self._h, name.encode(), C.byref(r), C.byref(c)))
        out = np.empty(c.value)
        check(self.L.bvhar_cuda_fit_record_mean(self._h, name.encode(), out.ctypes.data_as(c_dp), c.value))
        return out

    SUMMARY_COLUMNS = ("mean", "sd", "lower", "median", "upper", "rhat", "ess", "pip")

    def record_summary(self, name: str, level: float = .05) -> Dict[str, np.ndarray]:
        """Posterior summaries of each column over all chains and kept draws, reduced on the device
        (``bvhar_cuda_fit_record_summary``): mean, sd, the level / 2, 1 / 2 and 1 - level / 2 quantiles, split-R-hat, ESS and the
        share of non-zero entries (the PIP of the SAVS records)."""
        r, c = C.c_int64(), C.c_int64()
        check(self.L.bvhar_cuda_fit_record_dims(self._h, name.encode(), C.byref(r), C.byref(c)))
        out = np.empty((c.value, 8))
        check(self.L.bvhar_cuda_fit_record_summary(self._h, name.encode(), float(level), out.ctypes.data_as(c_dp), c.value))
        return {nm: out[:, i].copy() for i, nm in enumerate(self.SUMMARY_COLUMNS)}

    def records(self, window: int = 0, chain: int = 0) -> Dict[str, np.ndarray]:
        return {nm: self.record(nm, window, chain) for nm in self.record_names()}


class _McmcRun:
    """``bvhar::McmcRun<BaseMcmc, isGroup>`` (triangular.h:1194-1233)."""

    _is_group = True
    _sv: Optional[bool] = None

    def __init__(self, num_chains, num_iter, num_burn, thin, x, y, param_cov, param_prior, param_intercept,
                 param_init, prior_type, grp_id, own_id, cross_id, grp_mat, include_mean, seed_chain,
                 display_progress=False, nthreads=1, device: int = 0, record_mask: int = REC_ALL):
        is_sv = "initial_mean" in param_cov
        if self._sv is not None and is_sv != self._sv:
            raise ValueError("param_cov does not match the model class (initial_mean selects stochastic volatility)")
        self._display = bool(display_progress)
        self._packed = PackedFit(num_chains, num_iter, num_burn, thin, x, y, param_cov, param_prior, param_intercept,
                                 param_init, prior_type, grp_id, own_id, cross_id, grp_mat, include_mean, seed_chain,
                                 is_group=self._is_group, device=device, record_mask=record_mask)
        self._fit = CudaFit(self._packed)
        self._num_iter = int(num_iter)

    def returnRecords(self) -> List[Dict[str, np.ndarray]]:
        def progress(done, total):  # 5 % messages of triangular.h:1255-1274
            if self._display:
                print(f"[all chains] {done} / {total}", flush=True)
            return False

        try:
            self._fit.run(progress if self._display else None)
        except BvharCudaError as e:
            if e.status != 3:  # numerical flags do not abort; everything else does
                raise
        return [self._fit.records(0, c) for c in range(self._packed.n_chains)]


class McmcLdlt(_McmcRun):
    _is_group, _sv = True, False


class McmcLdltGrp(_McmcRun):
    _is_group, _sv = False, False


class SvMcmc(_McmcRun):
    _is_group, _sv = True, True


class SvGrpMcmc(_McmcRun):
    _is_group, _sv = False, True


# ------------------------------------------------------------------------------------------------
def _stable_rows(coef: np.ndarray, dim: int, nrow: int, har: Optional[np.ndarray], device: int = 0, threshold: float = 1.0) -> np.ndarray:
    """``is_stable`` per draw (draw.h:1265-1295: companion-matrix spectral radius < threshold), decided on the device by
    ``bvhar_cuda_is_stable`` (repeated squaring of the companion matrix with the DMMA contraction, kernels_stable.cu).
    ``har``: the top-left 3 dim x (month dim) corner of the HAR transformation for VHAR draws (forecaster.h:319)."""
    from ._abi import load_library
    draws = np.ascontiguousarray(coef, dtype=np.float64)
    n = draws.shape[0]
    out = np.zeros(n, dtype=np.int32)
    if n == 0:
        return out.astype(bool)
    if har is not None:
        hb = np.asfortranarray(har, dtype=np.float64)
        lag = hb.shape[1] // dim
        hp, ldh = hb.ctypes.data_as(c_dp), hb.shape[0]
    else:
        lag = nrow // dim
        hp, ldh = None, 0
    check(load_library().bvhar_cuda_is_stable(int(device), n, dim, nrow, lag, draws.ctypes.data_as(c_dp), draws.shape[1], hp, ldh,
                                              float(threshold), out.ctypes.data_as(C.POINTER(C.c_int32))))
    return out.astype(bool)


def spillover(coef: np.ndarray, contem: np.ndarray, sv: np.ndarray, dim: int, nrow: int, lag: int, step: int,
              har_trans: Optional[np.ndarray] = None, sv_kind: int = 0, n_time: int = 0, device: int = 0) -> Dict[str, np.ndarray]:
    """``McmcSpilloverRun<RecordType>::returnSpillover()`` (spillover.h:246-260) for the draws in the rows of ``coef`` /
    ``contem`` / ``sv``: the ``connect`` / ``to`` / ``from`` / ``tot`` / ``net`` / ``net_pairwise`` densities, computed on
    the device (``bvhar_cuda_spillover``).  ``sv_kind``: 0 = ``d_record`` rows, 1 = log-volatilities of one time point,
    2 = whole ``h_record`` rows (``n_time`` x dim, time average of exp(h / 2), config.h:980-988)."""
    from ._abi import SpilloverDesc, load_library
    cf = np.ascontiguousarray(coef, dtype=np.float64)
    S = cf.shape[0]
    q = dim * (dim - 1) // 2
    ac = np.ascontiguousarray(contem, dtype=np.float64).reshape(S, -1) if q > 0 else np.zeros((S, 1))
    sc = np.ascontiguousarray(sv, dtype=np.float64).reshape(S, -1)
    D = SpilloverDesc()
    D.abi_version = _abi.ABI_VERSION; D.device = int(device); D.dim = dim; D.nrow = nrow; D.lag = lag; D.step = int(step)
    D.sv_kind = int(sv_kind); D.n_time = int(n_time); D.n_draws = S
    D.coef = cf.ctypes.data_as(c_dp); D.ld_coef = cf.shape[1]
    D.contem = ac.ctypes.data_as(c_dp); D.ld_contem = ac.shape[1]
    D.sv = sc.ctypes.data_as(c_dp); D.ld_sv = sc.shape[1]
    hb = None
    if har_trans is not None:
        hb = np.asfortranarray(np.asarray(har_trans, dtype=np.float64)[: 3 * dim, : lag * dim])
        D.har_trans = hb.ctypes.data_as(c_dp); D.ldh = hb.shape[0]
    So = S * int(n_time) if int(sv_kind) == 3 else S  # sv_kind 3: one spillover per time point, ordered [time][draw]
    connect = np.empty((dim, So * dim), order="F"); net_pw = np.empty((dim, So * dim), order="F")
    to = np.empty(So * dim); frm = np.empty(So * dim); tot = np.empty(So)
    check(load_library().bvhar_cuda_spillover(C.byref(D), connect.ctypes.data_as(c_dp), to.ctypes.data_as(c_dp), frm.ctypes.data_as(c_dp),
                                              tot.ctypes.data_as(c_dp), net_pw.ctypes.data_as(c_dp)))
    return {"connect": connect, "to": to, "from": frm, "tot": tot, "net": to - frm, "net_pairwise": net_pw}


def dynamic_spillover(packed: PackedFit, step: int, var_lag: int, har_trans: Optional[np.ndarray] = None, sparse: bool = False):
    """``DynamicLdltSpillover::returnSpillover()`` (spillover.h:272-443): every (window, chain) of ``packed``'s window table
    is refitted and its spillover computed from the device-resident draws (``bvhar_cuda_dynamic_spillover``).
    Returns ``to, from [window, chain, draw, variable]`` and ``tot [window, chain, draw]``."""
    from ._abi import DynspillParams, load_library
    P = DynspillParams()
    P.abi_version = _abi.ABI_VERSION; P.step = int(step); P.var_lag = int(var_lag); P.sparse = int(bool(sparse))
    hb = None
    if har_trans is not None:
        hb = np.asfortranarray(np.asarray(har_trans, dtype=np.float64)[: 3 * packed.dim, : var_lag * packed.dim])
        P.har_trans = hb.ctypes.data_as(c_dp); P.ldh = hb.shape[0]
    d = packed.desc
    thin = max(1, int(d.thin))
    S = (max(0, int(d.num_iter) - int(d.num_burn)) + thin - 1) // thin
    W, Cn, k = packed.n_windows, packed.n_chains, packed.dim
    to = np.empty((W, Cn, S, k)); frm = np.empty((W, Cn, S, k)); tot = np.empty((W, Cn, S))
    check(load_library().bvhar_cuda_dynamic_spillover(C.byref(d), C.byref(P), to.ctypes.data_as(c_dp), frm.ctypes.data_as(c_dp),
                                                      tot.ctypes.data_as(c_dp)))
    return to, frm, tot


def pack_forecast(units: Sequence[Dict[str, np.ndarray]], last_obs: Sequence[np.ndarray], step: int, var_lag: int,
                  include_mean: bool, sv_model: bool, seeds, har_trans: Optional[np.ndarray] = None, sv: bool = True,
                  valid_vec: Optional[np.ndarray] = None, stable: bool = False, sparse: bool = False,
                  coef_name: str = "alpha", device: int = 0, unit_id_base: int = 0, level: float = 0.0):
    """Build the ``bvhar_forecast_desc`` of a batch of forecasters from their record dicts
    (``initialize_record`` config.h:1316-1363 + the forecaster ctors forecaster.h:27-52, 229-240).
    Returns (descriptor, keep-alive list)."""
    if sparse and level > 0:
        raise ValueError("If 'level > 0', 'spare' should be false.")  # forecaster.h:445-447 (sic)
    dim = last_obs[0].shape[1]
    sfx = "_sparse_record" if sparse else "_record"
    coefs, contems, diags, sighs = [], [], [], []
    nrow = None
    for rec in units:
        a = np.asarray(rec[coef_name + sfx], dtype=np.float64)
        nrow = a.shape[1] // dim
        full = np.hstack([a, np.asarray(rec["c" + sfx])]) if include_mean else a
        sel = slice(None)
        if stable:  # subsetStable, config.h:884-914, 1003-1034
            hb = None if har_trans is None else har_trans[: 3 * dim, : var_lag * dim]
            sel = _stable_rows(full, dim, nrow, hb)
        coefs.append(full[sel]); contems.append(np.asarray(rec["a" + sfx])[sel])
        if sv_model:
            h = rec["h_last_record"] if "h_last_record" in rec else np.asarray(rec["h_record"])[:, -dim:]
            diags.append(np.asarray(h)[sel]); sighs.append(np.asarray(rec["sigh_record"])[sel])
        else:
            diags.append(np.asarray(rec["d_record"])[sel])
    S = coefs[0].shape[0]
    if any(c.shape[0] != S for c in coefs):
        raise ValueError("all units must keep the same number of draws")
    if S == 0:
        raise ValueError("No stable MCMC draws")
    U = len(units)
    f = lambda lst: np.ascontiguousarray(np.concatenate([np.asfortranarray(m).reshape(-1, order="F") for m in lst]))
    coef_b, contem_b, diag_b = f(coefs), f(contems), f(diags)
    sigh_b = f(sighs) if sv_model else None
    last_b = f([np.asarray(lo, dtype=np.float64)[-var_lag:, :] for lo in last_obs])
    D = ForecastDesc()
    D.abi_version = _abi.ABI_VERSION; D.device = device; D.n_units = U; D.dim = dim; D.nrow_coef = nrow
    D.include_mean = int(include_mean); D.cov_type = _abi.COV_SV if sv_model else _abi.COV_LDLT; D.step = int(step)
    D.var_lag = int(var_lag); D.is_vhar = int(har_trans is not None); D.sv = int(sv); D.get_lpl = int(valid_vec is not None)
    D.n_draws = S; D.unit_id_base = int(unit_id_base); D.level = float(level)
    keep = [coef_b, contem_b, diag_b, last_b]
    D.coef_record = coef_b.ctypes.data_as(c_dp); D.contem_record = contem_b.ctypes.data_as(c_dp)
    D.diag_record = diag_b.ctypes.data_as(c_dp); D.last_obs = last_b.ctypes.data_as(c_dp)
    if sv_model:
        keep.append(sigh_b); D.sigh_record = sigh_b.ctypes.data_as(c_dp)
    if har_trans is not None:
        ht = np.asfortranarray(har_trans, dtype=np.float64); keep.append(ht); D.har_trans = ht.ctypes.data_as(c_dp)
    if valid_vec is not None:
        vv = np.ascontiguousarray(valid_vec, dtype=np.float64); keep.append(vv); D.valid_vec = vv.ctypes.data_as(c_dp)
    sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32); keep.append(sd)
    D.seed = sd.ctypes.data_as(c_up)
    return D, keep


def forecast_density(units: Sequence[Dict[str, np.ndarray]], last_obs: Sequence[np.ndarray], step: int, var_lag: int,
                     include_mean: bool, sv_model: bool, seeds, har_trans: Optional[np.ndarray] = None, sv: bool = True,
                     valid_vec: Optional[np.ndarray] = None, stable: bool = False, sparse: bool = False,
                     coef_name: str = "alpha", device: int = 0, unit_id_base: int = 0, level: float = 0.0):
    """Density forecasts of independent units on the GPU (``McmcForecaster::forecastDensity``,
    forecaster.h:55-85).  ``level > 0``: the Select forecasters (forecaster.h:348-416), coefficients masked by each unit's
    credible-interval activity graph.  Returns (list of step x (S*k) matrices, step-wise ALPL per unit or None)."""
    L = load_library()
    D, keep = pack_forecast(units, last_obs, step, var_lag, include_mean, sv_model, seeds, har_trans, sv, valid_vec, stable,
                            sparse, coef_name, device, unit_id_base, level)
    U, S, dim = D.n_units, D.n_draws, D.dim
    out = np.empty(U * step * S * dim)
    lpl = np.zeros(U * step) if valid_vec is not None else None
    check(L.bvhar_cuda_forecast_density(C.byref(D), out.ctypes.data_as(c_dp), lpl.ctypes.data_as(c_dp) if lpl is not None else c_dp()))
    dens = [out[u * step * S * dim:(u + 1) * step * S * dim].reshape(step, S * dim, order="F") for u in range(U)]
    return dens, (lpl.reshape(U, step) if lpl is not None else None)


def outforecast(packed: PackedFit, y_full: np.ndarray, step: int, var_lag: int, seed_forecast, har_trans=None, sv: bool = True,
                valid_vec=None, sparse: bool = False, level: float = 0.0):
    """``McmcOutforecastRun::returnForecast`` for the windows of ``packed`` in ONE native call
    (``bvhar_cuda_outforecast_run``): refit every (window, chain), forecast from the device-resident draws.
    Returns (forecast[window, chain, draw, variable] of the last step, lpl[window, chain] or None)."""
    L = load_library()
    P = _abi.OutforecastParams()
    P.abi_version = _abi.ABI_VERSION; P.step = int(step); P.var_lag = int(var_lag); P.is_vhar = int(har_trans is not None)
    P.sv = int(sv); P.get_lpl = int(valid_vec is not None); P.sparse = int(sparse); P.level = float(level)
    yf = np.asfortranarray(y_full, dtype=np.float64)
    P.y_full = yf.ctypes.data_as(c_dp); P.y_rows = yf.shape[0]
    keep = [yf]
    if har_trans is not None:
        ht = np.asfortranarray(har_trans, dtype=np.float64); keep.append(ht); P.har_trans = ht.ctypes.data_as(c_dp)
    if valid_vec is not None:
        vv = np.ascontiguousarray(valid_vec, dtype=np.float64); keep.append(vv); P.valid_vec = vv.ctypes.data_as(c_dp)
    sd = np.ascontiguousarray(np.asarray(seed_forecast, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32); keep.append(sd)
    P.seed_forecast = sd.ctypes.data_as(c_up)
    W, Cn, k = packed.n_windows, packed.n_chains, packed.dim
    d = packed.desc
    S = (d.num_iter - d.num_burn + d.thin - 1) // d.thin
    out = np.empty((W, Cn, S, k))
    lpl = np.zeros((W, Cn)) if valid_vec is not None else None
    check(L.bvhar_cuda_outforecast_run(C.byref(packed.desc), C.byref(P), out.ctypes.data_as(c_dp),
                                       lpl.ctypes.data_as(c_dp) if lpl is not None else c_dp()))
    return out, lpl
