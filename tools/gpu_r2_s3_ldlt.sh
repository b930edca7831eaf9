#!/bin/bash
# LDLT path on the Z'Z contraction + quadratic-form state + GDP log-product: GPU suite, then C5 / C1 per-kernel timings
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
(timeout 300 python tools/gpu_bench_cfg.py C5 256; timeout 300 python tools/gpu_bench_cfg.py C1 1024) 2>&1 | tee gpurun_out/s3_ldlt.log
