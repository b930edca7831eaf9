#!/bin/bash
# quick perf loop: per-kernel timings at C3 + the SV / trajectory parity tests
python tools/gpu_bench_cfg.py C3 1024 10 2>&1 | tail -22
python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_parity.py -m gpu -q -x 2>&1 | tail -4
