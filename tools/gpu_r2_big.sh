#!/bin/bash
# large-design path: parity at the BASELINE shapes, then per-kernel timing of C4 / C5
timeout 900 python -m pytest tests/test_gpu_baseline_parity.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
timeout 300 python tools/gpu_bench_cfg.py C4 128 2>&1 | tail -16
BVHAR_NO_DIRECT_P=1 timeout 300 python tools/gpu_bench_cfg.py C4 128 2>&1 | tail -16
timeout 300 python tools/gpu_bench_cfg.py C5 256 2>&1 | tail -12
