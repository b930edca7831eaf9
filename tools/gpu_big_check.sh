#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python tools/gpu_bench_cfg.py C3 1024 10 2>&1 | head -8
python tools/gpu_bench_cfg.py C4 1024 3 2>&1 | head -5
python tools/gpu_bench_cfg.py C5 256 3 2>&1 | head -4
python tools/gpu_bench_cfg.py C1 1024 10 2>&1 | head -5
