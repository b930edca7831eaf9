#!/bin/bash
# new persistent bulk-copy contraction: parity on the building-block shapes, then A/B timing against the cp.async kernel
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "dgemm" 2>&1 | tail -3
timeout 120 python tools/gpu_gemm_ab.py 2>&1 | tail -6
BVHAR_GEMM_LEGACY=1 timeout 120 python tools/gpu_gemm_ab.py 2>&1 | tail -6
