"""ctypes binding of the CPU oracle (oracle/bvhar_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product package bvhar_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(ORACLE_DIR)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from bvhar_b200._abi import FitDesc, ForecastDesc, PackedFit, c_dp  # noqa: E402  (interface structs only)
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libbvhar_oracle.so")
_lib = None


def build():
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("bvhar_oracle.cpp", "oracle_rng.h", "oracle_linalg.h")]
    srcs.append(os.path.join(ROOT, "include", "bvhar_cuda.h"))
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.oracle_fit_create.restype = vp
        L.oracle_fit_create.argtypes = [C.POINTER(FitDesc), C.c_int, C.c_int]
        L.oracle_fit_destroy.argtypes = [vp]
        L.oracle_fit_destroy.restype = None
        L.oracle_fit_set_threads.argtypes = [vp, C.c_int]
        L.oracle_fit_step.argtypes = [vp, C.c_int, C.c_int]
        L.oracle_fit_run.argtypes = [vp]
        L.oracle_fit_run_stage.argtypes = [vp, C.c_int, C.c_int]
        L.oracle_fit_state_len.argtypes = [vp, C.c_char_p]
        L.oracle_fit_state_len.restype = C.c_int64
        L.oracle_fit_get_state.argtypes = [vp, C.c_int, C.c_int, C.c_char_p, c_dp, C.c_int64]
        L.oracle_fit_set_state.argtypes = [vp, C.c_int, C.c_int, C.c_char_p, c_dp, C.c_int64]
        L.oracle_fit_set_sweep_count.argtypes = [vp, C.c_int]
        L.oracle_fit_record_dims.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.oracle_fit_record.argtypes = [vp, C.c_int, C.c_int, C.c_char_p, c_dp, C.c_int64, C.c_int64]
        L.oracle_fit_record_names.argtypes = [vp, C.c_char_p, C.c_int64]
        L.oracle_forecast_density.argtypes = [C.POINTER(ForecastDesc), C.c_int, c_dp, c_dp]
        L.oracle_draw_coef.argtypes = [C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
        L.oracle_draw_coef.restype = None
        L.oracle_draw_coef_packed.argtypes = [C.c_int, c_dp, c_dp, c_dp, c_dp]
        L.oracle_draw_coef_packed.restype = None
        L.oracle_dgemm_tn.argtypes = [C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp]
        L.oracle_dgemm_tn.restype = None
        L.oracle_spillover.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
        L.oracle_spillover.restype = None
        L.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.oracle_philox4x32_10.restype = None
        L.oracle_sample.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_int, c_dp]
        L.oracle_sample.restype = None
        L.oracle_draw_at.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_double]
        L.oracle_draw_at.restype = C.c_double
        L.oracle_minnesota_moments.argtypes = [C.c_int, C.c_int, c_dp, C.c_double, C.c_double, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
        L.oracle_minnesota_moments.restype = None
        L.oracle_mniw_run.argtypes = [C.c_int] * 6 + [c_dp, c_dp, c_dp, C.c_double, C.POINTER(C.c_uint32), C.c_uint32, C.c_int,
                                      C.c_double, C.c_double, C.c_double, C.c_double, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                      C.POINTER(C.c_uint8)]
        L.oracle_max_threads.restype = C.c_int
        L.oracle_use_blas.argtypes = [C.c_char_p, C.c_char_p]
        _lib = L
    return _lib


def dp(a):
    return a.ctypes.data_as(c_dp)


def use_scipy_openblas(on: bool = True) -> str:
    """Route the oracle's large dense products / Cholesky factorisations through the OpenBLAS bundled with SciPy
    (single-threaded per call; bench.py's reference arm only).  Returns a description of the back end in use."""
    if not on:
        lib().oracle_use_blas(None, None)
        return "hand-written loops, no BLAS"
    import glob
    import scipy
    cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so"))
    for path in cands:
        if lib().oracle_use_blas(path.encode(), b"scipy_") == 0:
            return f"OpenBLAS bundled with SciPy {scipy.__version__} ({os.path.basename(path)}: dgemm / dpotrf, 1 thread per call)"
    return "hand-written loops, no BLAS (no bundled OpenBLAS found)"


class OracleFit:
    """Same surface as bvhar_b200._engine.CudaFit, backed by the CPU oracle."""

    def __init__(self, packed: PackedFit, faithful: bool = False, rng: str = "philox", threads: int = 0):
        self.packed = packed
        self.L = lib()
        self.h = self.L.oracle_fit_create(C.byref(packed.desc), int(faithful), 0 if rng == "philox" else 1)
        if threads:
            self.L.oracle_fit_set_threads(self.h, threads)

    def close(self):
        if self.h:
            self.L.oracle_fit_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def step(self, n=1, record=False):
        self.L.oracle_fit_step(self.h, n, int(record))

    def run(self):
        self.L.oracle_fit_run(self.h)

    def run_stage(self, stage, it):
        self.L.oracle_fit_run_stage(self.h, stage, it)

    def set_sweep_count(self, c):
        self.L.oracle_fit_set_sweep_count(self.h, c)

    def get_state(self, name, window=0, chain=0):
        n = self.L.oracle_fit_state_len(self.h, name.encode())
        if name in ("latent_innov", "lvol_draw", "sqrt_sv"):  # n x k of THIS window (ragged window tables)
            n = int(self.packed.win_nrow[window]) * self.packed.dim
        if n == 0:  # e.g. the contemporaneous block of a univariate model
            return np.empty(0)
        out = np.empty(n)
        rc = self.L.oracle_fit_get_state(self.h, window, chain, name.encode(), dp(out), n)
        if rc:
            raise KeyError(name)
        return out

    def set_state(self, name, value, window=0, chain=0):
        v = np.ascontiguousarray(np.asarray(value, dtype=np.float64).reshape(-1, order="F"))
        rc = self.L.oracle_fit_set_state(self.h, window, chain, name.encode(), dp(v), v.size)
        if rc:
            raise KeyError(name)

    def record_names(self):
        buf = C.create_string_buffer(4096)
        self.L.oracle_fit_record_names(self.h, buf, 4096)
        return [s for s in buf.value.decode().split("\n") if s]

    def record(self, name, window=0, chain=0):
        r, c = C.c_int64(), C.c_int64()
        if self.L.oracle_fit_record_dims(self.h, name.encode(), C.byref(r), C.byref(c)):
            raise KeyError(name)
        out = np.empty((r.value, c.value), order="F")
        rc = self.L.oracle_fit_record(self.h, window, chain, name.encode(), dp(out), r.value, c.value)
        assert rc == 0, rc
        return out


def mniw_run(num_chains, num_iter, num_burn, thin, mn_mean, mn_prec, iw_scale, iw_shape, seed_chain, mh=None, unit_base=0):
    """oracle_mniw_run: McmcMniw / MhMinnesota (minnesota.h:241-278, 413-510) on the addressed Philox stream.  ``mh`` =
    dict(par [chains][1+k], chol_var [chains] upper factors of scale_variance * hessian^-1, lambda_param, psi_param).
    Returns (status, records) with records laid out like bvhar_b200.minnesota (rows x cols per chain)."""
    mean = np.asfortranarray(mn_mean, dtype=np.float64); prec = np.asfortranarray(mn_prec, dtype=np.float64)
    scale = np.asfortranarray(iw_scale, dtype=np.float64)
    d, k = mean.shape
    thin = max(int(thin), 1)
    rows = (num_iter - num_burn + thin - 1) // thin
    seeds = np.ascontiguousarray(seed_chain, dtype=np.uint32)
    alpha = np.zeros((num_chains, d * k, rows)); sigma = np.zeros((num_chains, k * k, rows))
    lam = np.zeros((num_chains, rows)); psi = np.zeros((num_chains, k, rows)); acc = np.zeros((num_chains, rows), dtype=np.uint8)
    par = np.zeros((num_chains, 1 + k)); G = np.zeros((num_chains, 1 + k, 1 + k)); lp = (1.0, 1.0); pp = (1.0, 1.0)
    if mh is not None:
        par = np.ascontiguousarray(mh["par"], dtype=np.float64); G = np.ascontiguousarray(mh["chol_var"], dtype=np.float64)
        lp = mh["lambda_param"]; pp = mh["psi_param"]
    rc = lib().oracle_mniw_run(num_chains, num_iter, num_burn, thin, k, d, dp(mean), dp(prec), dp(scale), float(iw_shape),
                               seeds.ctypes.data_as(C.POINTER(C.c_uint32)), unit_base, 1 if mh is not None else 0,
                               float(lp[0]), float(lp[1]), float(pp[0]), float(pp[1]), dp(par), dp(G), dp(alpha), dp(sigma),
                               dp(lam), dp(psi), acc.ctypes.data_as(C.POINTER(C.c_uint8)))
    out = []
    for c in range(num_chains):
        o = {"alpha_record": alpha[c].T, "sigma_record": sigma[c].T}
        if mh is not None:
            o.update(lambda_record=lam[c].copy(), psi_record=psi[c].T, accept_record=acc[c].astype(bool))
        out.append(o)
    return rc, out
