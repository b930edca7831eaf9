"""One-off: is the contraction's result finite and close to NumPy's for the shapes of tools/gpu_gemm_stress.py?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bvhar_b200 import _abi
L = _abi.load_library()
rng = np.random.default_rng(0)
for (M, N, K) in ((1000, 25600, 301), (1000, 25600, 61), (2560, 61, 1000), (20480, 61, 1000), (256, 33, 500), (128, 64, 100), (1000, 61, 300)):
    A = rng.normal(size=(K, M)); B = rng.normal(size=(K, N))
    want = A.T @ B
    C = np.full((M, N), 7.0)
    _abi.check(L.bvhar_cuda_dgemm_tn(0, M, N, K, A.ctypes.data_as(_abi.c_dp), M, B.ctypes.data_as(_abi.c_dp), N, C.ctypes.data_as(_abi.c_dp), N))
    err = np.abs(C - want)
    print(f"M={M} N={N} K={K}: nan {int(np.isnan(C).sum())}  max err {np.nanmax(err):.3e}  untouched {int((C == 7.0).sum())}", flush=True)
