"""Repeated parity of the persistent contraction on shapes with partial row tiles, ragged k-tiles and many tiles per CTA."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bvhar_b200 import _abi
L = _abi.load_library()
rng = np.random.default_rng(0)
for (M, N, K) in ((1000, 25600, 301), (1000, 3000, 301), (200, 20301, 400), (72, 20301, 397), (2560, 1891, 1000), (1000, 25600, 61)):
    A = rng.normal(size=(K, M)); B = rng.normal(size=(K, N))
    want = A.T @ B
    worst = 0.0
    for rep in range(6):
        C = np.zeros((M, N))
        _abi.check(L.bvhar_cuda_dgemm_tn(0, M, N, K, A.ctypes.data_as(_abi.c_dp), M, B.ctypes.data_as(_abi.c_dp), N, C.ctypes.data_as(_abi.c_dp), N))
        err = np.abs(C - want).max()
        worst = max(worst, err)
    print(f"M={M} N={N} K={K}: worst abs err over 6 runs {worst:.3e}", flush=True)
