#!/bin/bash
# batched large-design Cholesky: parity at the BASELINE shapes + forced big path, then C4 / C5 timings (batched vs CTA-per-matrix)
timeout 900 python -m pytest tests/test_gpu_baseline_parity.py tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q 2>&1 | tail -3
for c in "C4 128" "C4 1024 3" "C5 256"; do timeout 300 python tools/gpu_bench_cfg.py $c 2>&1 | grep -E "^C|graph|chol|seq_big"; done
for c in "C4 128" "C5 256"; do BVHAR_CHOL_CTA=1 timeout 300 python tools/gpu_bench_cfg.py $c 2>&1 | grep -E "graph|chol"; done
