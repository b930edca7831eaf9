"""G1-shaped contraction timed alone: python tools/gpu_gemm_ab.py  (BVHAR_GEMM_LEGACY=1 selects the cp.async kernel)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bvhar_b200 import _abi
L = _abi.load_library()
tag = "legacy cp.async" if os.environ.get("BVHAR_GEMM_LEGACY") == "1" else "persistent bulk/mbarrier"
for (M, N, K, what) in ((20480, 1891, 1000, "C3 G1 (1 024 chains)"), (20480, 61, 1000, "C3 G2"), (8192, 8192, 8192, "square 8192"),
                        (25600, 20301, 400, "C4-like full tiles"), (2560, 1891, 1000, "C3 G1 at 128 chains")):
    ms = C.c_double()
    _abi.check(L.bvhar_cuda_bench_dgemm_tn(0, M, N, K, 10, C.byref(ms)))
    print(f"{tag}: {what}: M={M} N={N} K={K}: {ms.value:.4f} ms  {2.0 * M * N * K / ms.value / 1e9:.2f} TFLOP/s", flush=True)
