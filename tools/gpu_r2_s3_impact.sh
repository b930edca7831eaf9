#!/bin/bash
# new impact_sv inner loop + LDLT small-k path: GPU suite, then per-kernel timings of every configuration
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for c in "C3 1024" "C1 1024" "C2 64" "C4 128" "C5 256"; do timeout 300 python tools/gpu_bench_cfg.py $c; done 2>&1 | tee gpurun_out/s3_impact.log
