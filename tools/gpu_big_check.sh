#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q 2>&1 | tail -3
python tools/gpu_bench_cfg.py C3 1024 10 2>&1 | head -18
python tools/gpu_bench_cfg.py C2 64 10 2>&1 | head -5
