"""ctypes mirror of ``include/bvhar_cuda.h`` and the loader of the CUDA engine's shared library.

The product path has no CPU fallback: if ``libbvhar_b200.so`` is missing or was not built, every
entry point raises ``RuntimeError`` telling the user to run ``python -c "import __graft_entry__ as
g; g.build()"`` (or ``make -C bvhar_b200/csrc``).

The descriptor packing done here is what the reference's header adapter does when it unpacks the
``LIST`` arguments of ``McmcRun`` (inst/include/bvhar/src/bayes/triangular/triangular.h:1194-1210,
config.h:50-575) — same keys, same shapes.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

ABI_VERSION = 2

COV_LDLT, COV_SV = 0, 1
PRIOR_MINNESOTA, PRIOR_SSVS, PRIOR_HORSESHOE, PRIOR_HIERMINN, PRIOR_NG, PRIOR_DL, PRIOR_GDP = 1, 2, 3, 4, 5, 6, 7
REC_COEF, REC_SPARSE, REC_PRIOR, REC_LVOL, REC_LVOL_LAST = 1, 2, 4, 8, 16
REC_ALL = 0x0F

STATUS_NAMES = {0: "ok", 1: "bad argument", 2: "CUDA error", 3: "numerical failure", 4: "interrupted", 5: "unsupported"}

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_up = C.POINTER(C.c_uint32)


class PriorParams(C.Structure):
    _fields_ = [
        ("minn_prior_mean", c_dp), ("minn_prior_prec", c_dp),
        ("hmn_shape", C.c_double), ("hmn_rate", C.c_double), ("hmn_grid_size", C.c_int32),
        ("ssvs_coef_s1", c_dp), ("ssvs_coef_s2", c_dp),
        ("ssvs_chol_s1", C.c_double), ("ssvs_chol_s2", C.c_double),
        ("ssvs_coef_slab_shape", C.c_double), ("ssvs_coef_slab_scl", C.c_double),
        ("ssvs_chol_slab_shape", C.c_double), ("ssvs_chol_slab_scl", C.c_double),
        ("ssvs_coef_grid", C.c_int32), ("ssvs_chol_grid", C.c_int32),
        ("ng_shape_sd", C.c_double), ("ng_group_shape", C.c_double), ("ng_group_scale", C.c_double),
        ("ng_global_shape", C.c_double), ("ng_global_scale", C.c_double),
        ("ng_contem_global_shape", C.c_double), ("ng_contem_global_scale", C.c_double),
        ("dl_grid_size", C.c_int32), ("dl_shape", C.c_double), ("dl_scale", C.c_double),
        ("gdp_grid_shape", C.c_int32), ("gdp_grid_rate", C.c_int32),
    ]


class ChainInit(C.Structure):
    _fields_ = [
        ("init_coef", c_dp), ("init_contem", c_dp), ("init_diag", c_dp),
        ("lvol_init", c_dp), ("lvol", c_dp), ("lvol_sig", c_dp),
        ("own_lambda", C.c_double), ("cross_lambda", C.c_double), ("contem_lambda", C.c_double),
        ("coef_dummy", c_dp), ("coef_mixture", c_dp), ("chol_mixture", c_dp), ("coef_slab", c_dp), ("contem_slab", c_dp),
        ("coef_spike_scl", C.c_double), ("chol_spike_scl", C.c_double),
        ("local_sparsity", c_dp), ("global_sparsity", C.c_double), ("group_sparsity", c_dp),
        ("contem_local_sparsity", c_dp), ("contem_global_sparsity", C.c_double),
        ("local_shape", c_dp), ("contem_shape", C.c_double),
        ("group_rate", c_dp), ("contem_rate", c_dp),
        ("gamma_shape", C.c_double), ("gamma_rate", C.c_double),
        ("contem_gamma_shape", C.c_double), ("contem_gamma_rate", C.c_double),
    ]


class FitDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("n_windows", C.c_int32), ("n_chains", C.c_int32),
        ("dim", C.c_int32), ("dim_design", C.c_int32), ("include_mean", C.c_int32),
        ("cov_type", C.c_int32), ("prior_type", C.c_int32), ("is_group", C.c_int32),
        ("num_iter", C.c_int32), ("num_burn", C.c_int32), ("thin", C.c_int32),
        ("record_mask", C.c_uint32), ("lvol_from_init", C.c_int32), ("unit_id_base", C.c_int32),
        ("n_rows_total", C.c_int64),
        ("x", c_dp), ("y", c_dp), ("win_row0", c_ip), ("win_nrow", c_ip),
        ("grp_mat", c_ip), ("grp_id", c_ip), ("n_grp", C.c_int32), ("n_own", C.c_int32),
        ("own_id", c_ip), ("cross_id", c_ip), ("n_cross", C.c_int32), ("reserved1", C.c_int32),
        ("sig_shape", c_dp), ("sig_scale", c_dp), ("sv_init_mean", c_dp), ("sv_init_prec", c_dp),
        ("mean_non", c_dp), ("sd_non", C.c_double),
        ("prior", PriorParams),
        ("init", C.POINTER(ChainInit)),
        ("seed_chain", c_up),
    ]


class ForecastDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("n_units", C.c_int32), ("dim", C.c_int32),
        ("nrow_coef", C.c_int32), ("include_mean", C.c_int32), ("cov_type", C.c_int32), ("step", C.c_int32),
        ("var_lag", C.c_int32), ("is_vhar", C.c_int32), ("sv", C.c_int32), ("get_lpl", C.c_int32),
        ("n_draws", C.c_int32), ("unit_id_base", C.c_int32),
        ("har_trans", c_dp), ("coef_record", c_dp), ("contem_record", c_dp), ("diag_record", c_dp),
        ("sigh_record", c_dp), ("last_obs", c_dp), ("valid_vec", c_dp), ("seed", c_up), ("level", C.c_double),
    ]


class SpilloverDesc(C.Structure):
    """bvhar_spillover_desc (include/bvhar_cuda.h)."""
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("dim", C.c_int32), ("nrow", C.c_int32), ("lag", C.c_int32),
        ("step", C.c_int32), ("sv_kind", C.c_int32), ("n_time", C.c_int32), ("n_draws", C.c_int64),
        ("coef", c_dp), ("ld_coef", C.c_int64), ("contem", c_dp), ("ld_contem", C.c_int64), ("sv", c_dp), ("ld_sv", C.c_int64),
        ("har_trans", c_dp), ("ldh", C.c_int64),
    ]


class DynspillParams(C.Structure):
    """bvhar_dynspill_params (include/bvhar_cuda.h)."""
    _fields_ = [("abi_version", C.c_int32), ("step", C.c_int32), ("var_lag", C.c_int32), ("sparse", C.c_int32),
                ("har_trans", c_dp), ("ldh", C.c_int64)]


class OutforecastParams(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("step", C.c_int32), ("var_lag", C.c_int32), ("is_vhar", C.c_int32),
        ("sv", C.c_int32), ("get_lpl", C.c_int32), ("sparse", C.c_int32), ("reserved", C.c_int32),
        ("har_trans", c_dp), ("y_full", c_dp), ("y_rows", C.c_int64), ("valid_vec", c_dp), ("seed_forecast", c_up),
        ("level", C.c_double),
    ]


class MniwDesc(C.Structure):
    """bvhar_mniw_desc (include/bvhar_cuda.h): the conjugate Minnesota samplers."""
    _fields_ = [
        ("device", C.c_int32), ("num_chains", C.c_int32), ("num_iter", C.c_int32), ("num_burn", C.c_int32), ("thin", C.c_int32),
        ("dim", C.c_int32), ("dim_design", C.c_int32),
        ("mn_mean", c_dp), ("mn_prec", c_dp), ("iw_scale", c_dp), ("iw_shape", C.c_double),
        ("seed_chain", c_up), ("unit_id_base", C.c_uint32), ("mh", C.c_int32),
        ("gam_shape", C.c_double), ("gam_rate", C.c_double), ("invgam_shape", C.c_double), ("invgam_scl", C.c_double),
        ("init_par", c_dp), ("init_hess", c_dp), ("init_scale_variance", c_dp),
    ]


PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int)


def _f64(a, order="F") -> np.ndarray:
    if type(a) is np.ndarray and a.dtype == np.float64 and a.flags.aligned and (a.flags.f_contiguous if order == "F" else a.flags.c_contiguous):
        return a  # already in the layout the C-ABI reads (thousands of per-chain init vectors take this path)
    return np.require(np.asarray(a, dtype=np.float64), dtype=np.float64, requirements=["F" if order == "F" else "C", "A"])


def _i32(a) -> np.ndarray:
    return np.require(np.asarray(a, dtype=np.int32).reshape(-1, order="F"), dtype=np.int32, requirements=["C", "A"])


def _dp(a: Optional[np.ndarray]):
    return C.cast(a.__array_interface__["data"][0], c_dp) if a is not None else c_dp()


def _ip(a: Optional[np.ndarray]):
    return a.ctypes.data_as(c_ip) if a is not None else c_ip()


def _scalar(v) -> float:
    return float(np.asarray(v, dtype=np.float64).reshape(-1)[0])


class PackedFit:
    """Owns the numpy buffers a ``bvhar_fit_desc`` points into.

    Arguments follow ``bvhar::McmcRun`` (triangular.h:1194-1201) plus the window table of the
    out-of-sample runners (forecaster.h:851-875, 920-956).
    """

    def __init__(
        self, num_chains: int, num_iter: int, num_burn: int, thin: int, x, y,
        param_cov: Dict[str, Any], param_prior: Dict[str, Any], param_intercept: Dict[str, Any],
        param_init: Sequence[Dict[str, Any]], prior_type: int, grp_id, own_id, cross_id, grp_mat,
        include_mean: bool, seed_chain, is_group: bool = True, win_row0=None, win_nrow=None,
        record_mask: int = REC_ALL, device: int = 0, lvol_from_init: bool = False, unit_id_base: int = 0,
    ):
        self._keep: List[Any] = []
        k = self._keep.append
        X = _f64(x); Y = _f64(y)
        if X.ndim != 2 or Y.ndim != 2 or X.shape[0] != Y.shape[0]:
            raise ValueError("'x' and 'y' must be matrices with the same number of rows")
        n_tot, d = X.shape
        dim = Y.shape[1]
        q = dim * (dim - 1) // 2
        n_alpha = dim * d - (dim if include_mean else 0)
        sv = "initial_mean" in param_cov  # src/triangular.cpp:33
        if win_row0 is None:
            win_row0, win_nrow = [0], [n_tot]
        w0 = _i32(win_row0); wn = _i32(win_nrow)
        n_win = len(w0)
        if np.any(w0 < 0) or np.any(w0 + wn > n_tot) or np.any(wn <= 0):
            raise ValueError("window rows outside the design matrix")
        gm = _i32(grp_mat)
        if gm.size != n_alpha:
            raise ValueError(f"'grp_mat' must have {n_alpha} entries, got {gm.size}")
        gid = _i32(grp_id); oid = _i32(own_id); cid = _i32(cross_id)
        seeds = np.ascontiguousarray(np.asarray(seed_chain, dtype=np.int64).reshape(n_win, -1) & 0xFFFFFFFF, dtype=np.uint32)
        if seeds.shape != (n_win, num_chains):
            raise ValueError("'seed_chain' must be (n_windows x) num_chains")
        if len(param_init) != num_chains:
            raise ValueError("'param_init' needs one entry per chain")
        self.X, self.Y = X, Y
        for a in (X, Y, w0, wn, gm, gid, oid, cid, seeds):
            k(a)
        D = FitDesc()
        D.abi_version = ABI_VERSION; D.device = device
        D.n_windows = n_win; D.n_chains = num_chains; D.dim = dim; D.dim_design = d
        D.include_mean = int(bool(include_mean)); D.cov_type = COV_SV if sv else COV_LDLT
        D.prior_type = int(prior_type); D.is_group = int(bool(is_group))
        D.num_iter = int(num_iter); D.num_burn = int(num_burn); D.thin = int(thin)
        D.record_mask = int(record_mask); D.lvol_from_init = int(bool(lvol_from_init)); D.unit_id_base = int(unit_id_base)
        D.n_rows_total = n_tot
        D.x = _dp(X); D.y = _dp(Y); D.win_row0 = _ip(w0); D.win_nrow = _ip(wn)
        D.grp_mat = _ip(gm); D.grp_id = _ip(gid); D.n_grp = len(gid)
        D.own_id = _ip(oid); D.n_own = len(oid); D.cross_id = _ip(cid); D.n_cross = len(cid)

        def vec(dic, key, n, what):
            a = _f64(np.asarray(dic[key], dtype=np.float64).reshape(-1))
            if a.size == 1 and n > 1:
                a = _f64(np.repeat(a, n))
            if a.size != n:
                raise ValueError(f"'{what}[{key}]' must have length {n}, got {a.size}")
            k(a)
            return a

        D.sig_shape = _dp(vec(param_cov, "shape", dim, "param_cov"))
        D.sig_scale = _dp(vec(param_cov, "scale", dim, "param_cov"))
        if sv:
            D.sv_init_mean = _dp(vec(param_cov, "initial_mean", dim, "param_cov"))
            P0 = _f64(np.asarray(param_cov["initial_prec"], dtype=np.float64))
            if P0.size == 1:
                P0 = _f64(float(P0.reshape(-1)[0]) * np.eye(dim))
            if P0.shape != (dim, dim):
                raise ValueError("'param_cov[initial_prec]' must be dim x dim")
            k(P0); D.sv_init_prec = _dp(P0)
        if include_mean:
            D.mean_non = _dp(vec(param_intercept, "mean_non", dim, "param_intercept"))
            D.sd_non = _scalar(param_intercept["sd_non"])
        P = D.prior
        pt = int(prior_type)
        pp = param_prior
        if pt in (PRIOR_MINNESOTA, PRIOR_HIERMINN):
            from ._design import minnesota_moments  # host-side prior moments (config.h:124-151)
            mean, prec = minnesota_moments(pp, dim, n_alpha // dim, hierarchical=(pt == PRIOR_HIERMINN))
            mean = _f64(mean.reshape(-1, order="F")); prec = _f64(prec)
            k(mean); k(prec)
            P.minn_prior_mean = _dp(mean); P.minn_prior_prec = _dp(prec)
            if pt == PRIOR_HIERMINN:
                P.hmn_shape = _scalar(pp["shape"]); P.hmn_rate = _scalar(pp["rate"]); P.hmn_grid_size = int(pp["grid_size"])
        elif pt == PRIOR_SSVS:
            P.ssvs_coef_s1 = _dp(vec(pp, "coef_s1", len(gid), "param_prior"))
            P.ssvs_coef_s2 = _dp(vec(pp, "coef_s2", len(gid), "param_prior"))
            P.ssvs_chol_s1 = _scalar(pp["chol_s1"]); P.ssvs_chol_s2 = _scalar(pp["chol_s2"])
            P.ssvs_coef_slab_shape = _scalar(pp["coef_slab_shape"]); P.ssvs_coef_slab_scl = _scalar(pp["coef_slab_scl"])
            P.ssvs_chol_slab_shape = _scalar(pp["chol_slab_shape"]); P.ssvs_chol_slab_scl = _scalar(pp["chol_slab_scl"])
            P.ssvs_coef_grid = int(pp["coef_grid"]); P.ssvs_chol_grid = int(pp["chol_grid"])
        elif pt == PRIOR_NG:
            P.ng_shape_sd = _scalar(pp["shape_sd"])
            P.ng_group_shape = _scalar(pp["group_shape"]); P.ng_group_scale = _scalar(pp["group_scale"])
            P.ng_global_shape = _scalar(pp["global_shape"]); P.ng_global_scale = _scalar(pp["global_scale"])
            P.ng_contem_global_shape = _scalar(pp["contem_global_shape"]); P.ng_contem_global_scale = _scalar(pp["contem_global_scale"])
        elif pt == PRIOR_DL:
            P.dl_grid_size = int(pp["grid_size"]); P.dl_shape = _scalar(pp["shape"]); P.dl_scale = _scalar(pp["scale"])
        elif pt == PRIOR_GDP:
            P.gdp_grid_shape = int(pp["grid_shape"]); P.gdp_grid_rate = int(pp["grid_rate"])
        elif pt != PRIOR_HORSESHOE:
            raise ValueError(f"unknown prior_type {prior_type}")

        n0 = int(wn[0])
        inits = (ChainInit * num_chains)()
        # every field of bvhar_chain_init is 8 bytes wide (a pointer or a double): the table is filled column by column through
        # numpy views of its memory instead of ~10 ctypes attribute assignments and pointer casts per chain (thousands of chains)
        slots = C.sizeof(ChainInit) // 8
        raw_u = np.frombuffer(inits, dtype=np.uint64).reshape(num_chains, slots)
        raw_f = raw_u.view(np.float64)
        slot_of = {nm: getattr(ChainInit, nm).offset // 8 for nm, _ in ChainInit._fields_}
        pcols: Dict[str, List[int]] = {}
        scols: Dict[str, List[float]] = {}

        addressof, from_buffer = C.addressof, C.c_char.from_buffer

        def setp(field, arr):
            if arr is None:
                ptr = 0
            else:
                try:  # a third of the cost of arr.__array_interface__ (which builds a dict per call); needs a writable,
                    # C-contiguous buffer: the column-major matrices are passed as their transposed view (same memory)
                    ptr = addressof(from_buffer(arr if arr.ndim == 1 else arr.T))
                except (TypeError, ValueError):  # read-only or empty array
                    ptr = arr.__array_interface__["data"][0]
            col = pcols.get(field)
            if col is None:
                col = pcols[field] = []
            col.append(ptr)

        def sets(field, v):
            col = scols.get(field)
            if col is None:
                col = scols[field] = []
            col.append(v if type(v) is float else _scalar(v))

        f8 = np.float64
        for c, ini in enumerate(param_init):

            def ivec(key, n, required=True):
                if key not in ini:
                    if required:
                        raise ValueError(f"'param_init[{c}]' lacks '{key}'")
                    return None
                a = ini[key]
                if not (type(a) is np.ndarray and a.dtype == f8 and a.ndim == 1 and a.flags.c_contiguous and a.flags.aligned):
                    a = _f64(np.asarray(a, dtype=np.float64).reshape(-1, order="F"))
                if a.size == 1 and n > 1:
                    a = _f64(np.repeat(a, n))
                if a.size != n:
                    raise ValueError(f"'param_init[{c}][{key}]' must have {n} entries, got {a.size}")
                k(a)
                return a

            coef = _f64(np.asarray(ini["init_coef"], dtype=np.float64))
            if coef.shape != (d, dim):
                raise ValueError(f"'init_coef' must be {d} x {dim}")
            k(coef); setp("init_coef", coef)
            setp("init_contem", ivec("init_contem", q))
            if sv:
                setp("lvol_init", ivec("lvol_init", dim))
                setp("lvol_sig", ivec("lvol_sig", dim))
                if not lvol_from_init:
                    lv = _f64(np.asarray(ini["lvol"], dtype=np.float64))
                    if lv.shape != (n0, dim):
                        raise ValueError(f"'lvol' must be {n0} x {dim}")
                    if n_win > 1 and np.any(wn != n0):
                        raise ValueError("'lvol' inits need equal window lengths; use lvol_from_init")
                    k(lv); setp("lvol", lv)
            else:
                setp("init_diag", ivec("init_diag", dim))
            G = len(gid)
            if pt == PRIOR_HIERMINN:
                sets("own_lambda", ini["own_lambda"]); sets("cross_lambda", ini["cross_lambda"])
                sets("contem_lambda", ini["contem_lambda"])
            elif pt == PRIOR_SSVS:
                setp("coef_dummy", ivec("init_coef_dummy", n_alpha)); setp("coef_mixture", ivec("coef_mixture", G))
                setp("chol_mixture", ivec("chol_mixture", q)); setp("coef_slab", ivec("coef_slab", n_alpha))
                setp("contem_slab", ivec("contem_slab", q))
                sets("coef_spike_scl", ini["coef_spike_scl"]); sets("chol_spike_scl", ini["chol_spike_scl"])
            elif pt in (PRIOR_HORSESHOE, PRIOR_NG, PRIOR_DL):
                setp("local_sparsity", ivec("local_sparsity", n_alpha)); sets("global_sparsity", ini["global_sparsity"])
                setp("contem_local_sparsity", ivec("contem_local_sparsity", q))
                sets("contem_global_sparsity", ini["contem_global_sparsity"])
                if pt in (PRIOR_HORSESHOE, PRIOR_NG):
                    setp("group_sparsity", ivec("group_sparsity", G))
                if pt == PRIOR_NG:
                    setp("local_shape", ivec("local_shape", G)); sets("contem_shape", ini["contem_shape"])
            elif pt == PRIOR_GDP:
                setp("local_sparsity", ivec("local_sparsity", n_alpha)); setp("group_rate", ivec("group_rate", G))
                setp("contem_local_sparsity", ivec("contem_local_sparsity", q)); setp("contem_rate", ivec("contem_rate", q))
                sets("gamma_shape", ini["gamma_shape"]); sets("gamma_rate", ini["gamma_rate"])
                sets("contem_gamma_shape", ini["contem_gamma_shape"]); sets("contem_gamma_rate", ini["contem_gamma_rate"])
        for nm, col in pcols.items():
            raw_u[:, slot_of[nm]] = col
        for nm, col in scols.items():
            raw_f[:, slot_of[nm]] = col
        k(inits)
        D.init = C.cast(inits, C.POINTER(ChainInit))
        D.seed_chain = seeds.ctypes.data_as(c_up)
        self.desc = D
        self.n_windows, self.n_chains, self.dim, self.dim_design = n_win, num_chains, dim, d
        self.n_alpha, self.q, self.sv = n_alpha, q, sv
        self.win_nrow, self.win_row0, self.include_mean = wn, w0, bool(include_mean)


_LIB = None
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libbvhar_b200.so")


def lib_path() -> str:
    return _LIB_PATH


def load_library():
    """Load the CUDA engine; raise loudly if it has not been built (there is no CPU fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"bvhar_b200: CUDA engine not built ({_LIB_PATH} missing). Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C bvhar_b200/csrc`. "
            "There is no CPU fallback."
        )
    L = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    L.bvhar_cuda_last_error.restype = C.c_char_p
    L.bvhar_cuda_abi_version.restype = C.c_int
    L.bvhar_cuda_device_count.argtypes = [C.POINTER(C.c_int)]
    L.bvhar_cuda_struct_sizes.argtypes = [C.POINTER(C.c_int64)] * 4
    L.bvhar_cuda_host_alloc.argtypes = [C.c_int64, C.POINTER(vp)]
    L.bvhar_cuda_host_free.argtypes = [vp]
    L.bvhar_cuda_fit_create.argtypes = [C.POINTER(FitDesc), C.POINTER(vp)]
    L.bvhar_cuda_fit_run.argtypes = [vp, PROGRESS_FN, vp]
    L.bvhar_cuda_fit_step.argtypes = [vp, C.c_int, C.c_int]
    L.bvhar_cuda_fit_sync.argtypes = [vp]
    L.bvhar_cuda_fit_last_ms.argtypes = [vp, c_dp]
    L.bvhar_cuda_fit_set_profiling.argtypes = [vp, C.c_int]
    L.bvhar_cuda_fit_profile.argtypes = [vp, C.c_char_p, C.c_int64]
    L.bvhar_cuda_fit_launch_count.argtypes = [vp, C.POINTER(C.c_int64)]
    L.bvhar_cuda_fit_record_dims.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.bvhar_cuda_fit_record.argtypes = [vp, C.c_int, C.c_int, C.c_char_p, c_dp, C.c_int64, C.c_int64]
    L.bvhar_cuda_fit_record_all.argtypes = [vp, C.c_char_p, c_dp, C.c_int64, C.c_int64, C.c_int64]
    L.bvhar_cuda_fit_bind_record_sink.argtypes = [vp, C.c_char_p, c_dp, C.c_int64, C.c_int64, C.c_int64]
    L.bvhar_cuda_spillover.argtypes = [C.POINTER(SpilloverDesc), c_dp, c_dp, c_dp, c_dp, c_dp]
    L.bvhar_cuda_dynamic_spillover.argtypes = [C.POINTER(FitDesc), C.POINTER(DynspillParams), c_dp, c_dp, c_dp]
    L.bvhar_cuda_is_stable.argtypes = [C.c_int, C.c_int64, C.c_int32, C.c_int32, C.c_int32, c_dp, C.c_int64, c_dp, C.c_int64, C.c_double,
                                       C.POINTER(C.c_int32)]
    L.bvhar_cuda_fit_record_names.argtypes = [vp, C.c_char_p, C.c_int64]
    L.bvhar_cuda_fit_record_mean.argtypes = [vp, C.c_char_p, c_dp, C.c_int64]
    L.bvhar_cuda_fit_record_summary.argtypes = [vp, C.c_char_p, C.c_double, c_dp, C.c_int64]
    L.bvhar_cuda_fit_state_len.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64)]
    L.bvhar_cuda_fit_get_state.argtypes = [vp, C.c_int, C.c_int, C.c_char_p, c_dp, C.c_int64]
    L.bvhar_cuda_fit_set_state.argtypes = [vp, C.c_int, C.c_int, C.c_char_p, c_dp, C.c_int64]
    L.bvhar_cuda_fit_run_stage.argtypes = [vp, C.c_int, C.c_int]
    L.bvhar_cuda_fit_destroy.argtypes = [vp]
    L.bvhar_cuda_fit_destroy.restype = None
    L.bvhar_cuda_forecast_density.argtypes = [C.POINTER(ForecastDesc), c_dp, c_dp]
    L.bvhar_cuda_outforecast_run.argtypes = [C.POINTER(FitDesc), C.POINTER(OutforecastParams), c_dp, c_dp]
    L.bvhar_cuda_outforecast_params_size.argtypes = [C.POINTER(C.c_int64)]
    L.bvhar_cuda_dgemm_tn.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, c_dp, C.c_int64, c_dp, C.c_int64, c_dp, C.c_int64]
    L.bvhar_cuda_draw_coef_batched.argtypes = [C.c_int, C.c_int64, C.c_int32, c_dp, c_dp, c_dp, c_dp]
    L.bvhar_cuda_mniw_run.argtypes = [C.POINTER(MniwDesc), c_dp, c_dp, c_dp, c_dp, C.POINTER(C.c_uint8)]
    L.bvhar_cuda_bench_fp64_peak.argtypes = [C.c_int, C.c_int, c_dp]
    L.bvhar_cuda_bench_fp64_latency.argtypes = [C.c_int, C.c_int, c_dp]
    L.bvhar_cuda_bench_dgemm_tn.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, c_dp]
    if L.bvhar_cuda_abi_version() != ABI_VERSION:
        raise RuntimeError("bvhar_b200: libbvhar_b200.so was built against a different ABI version; rebuild it")
    sz = [C.c_int64() for _ in range(4)]
    L.bvhar_cuda_struct_sizes(*[C.byref(v) for v in sz])
    mine = [C.sizeof(FitDesc), C.sizeof(ChainInit), C.sizeof(ForecastDesc), C.sizeof(PriorParams)]
    if [v.value for v in sz] != mine:
        raise RuntimeError(f"bvhar_b200: ctypes struct layout {mine} differs from the library's {[v.value for v in sz]}")
    osz = C.c_int64()
    L.bvhar_cuda_outforecast_params_size(C.byref(osz))
    if osz.value != C.sizeof(OutforecastParams):
        raise RuntimeError("bvhar_b200: ctypes layout of bvhar_outforecast_params differs from the library's")
    _LIB = L
    return L


class BvharCudaError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"bvhar_cuda [{STATUS_NAMES.get(status, status)}]: {message}")
        self.status = status


def check(status: int):
    if status != 0:
        msg = load_library().bvhar_cuda_last_error()
        raise BvharCudaError(status, msg.decode() if msg else "")
