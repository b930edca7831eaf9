"""First GPU contact: fp64 peaks, contraction kernel check + timing, then stage parity."""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from bvhar_b200 import _abi
L = _abi.load_library()
v = C.c_double()
for mode, nm in ((0, "DFMA"), (1, "DMMA m8n8k4")):
    rc = L.bvhar_cuda_bench_fp64_peak(0, mode, C.byref(v)); print(nm, "peak TFLOP/s", v.value, "rc", rc, flush=True)
import oracle_binding as ob
rng = np.random.default_rng(0)
for (M, N, K) in ((37, 29, 50), (130, 257, 33), (256, 128, 64)):
    A = rng.normal(size=(K, M)); B = rng.normal(size=(K, N)); Cg = np.zeros((M, N)); Co = np.zeros((M, N))
    rc = L.bvhar_cuda_dgemm_tn(0, M, N, K, A.ctypes.data_as(_abi.c_dp), M, B.ctypes.data_as(_abi.c_dp), N, Cg.ctypes.data_as(_abi.c_dp), N)
    ref = A.T @ B
    print("gemm", M, N, K, "rc", rc, "max abs err", np.abs(Cg - ref).max(), flush=True)
for (M, N, K) in ((20480, 1891, 1000), (8192, 8192, 1024)):
    rc = L.bvhar_cuda_bench_dgemm_tn(0, M, N, K, 5, C.byref(v))
    print("gemm bench", M, N, K, "ms", v.value, "TFLOP/s", 2.0 * M * N * K / (v.value * 1e-3) / 1e12, flush=True)
import torch
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); b = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
torch.matmul(a, b); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): torch.matmul(a, b)
e1.record(); torch.cuda.synchronize()
print("cuBLAS dgemm 8192^3 TFLOP/s", 3 * 2 * 8192**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12, flush=True)
import gpu_check
for prior in ["Minnesota", "DL"]:
    for sv in (True, False):
        try:
            gpu_check.stage_check(prior, sv); gpu_check.sweep_check(prior, sv)
        except Exception:
            import traceback; traceback.print_exc()
