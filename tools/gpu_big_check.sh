#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -k "large or big" 2>&1 | tail -3
python tools/gpu_bench_cfg.py C4 1024 3 2>&1 | head -7
python tools/gpu_bench_cfg.py C5 256 3 2>&1 | head -6
