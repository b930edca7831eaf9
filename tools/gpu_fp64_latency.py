"""Dependent-issue latency of the fp64 operations (cycles per operation, one warp, one chain): python tools/gpu_fp64_latency.py"""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bvhar_b200 import _abi
L = _abi.load_library()
for kind, nm in enumerate(("DFMA", "DMMA m8n8k4 (accumulator chain)", "rsqrt(double) + add", "1.0 / x + add", "64-bit shuffle + DFMA", "DADD", "DMUL")):
    v = C.c_double()
    _abi.check(L.bvhar_cuda_bench_fp64_latency(0, kind, C.byref(v)))
    print(f"{nm:34s} {v.value:8.1f} cycles per dependent operation", flush=True)
