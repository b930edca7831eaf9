#!/bin/bash
# session-3 first pass: GPU suite, C5 reproducibility, large-design timings with / without the direct precision build, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/s3_gputests.log; cat gpurun_out/s3_gputests.log
timeout 300 python tools/gpu_c5_diag.py bench > gpurun_out/s3_c5_diag_bench.log 2>&1; grep -c "phase" gpurun_out/s3_c5_diag_bench.log; head -12 gpurun_out/s3_c5_diag_bench.log
timeout 300 python tools/gpu_c5_diag.py single > gpurun_out/s3_c5_diag_single.log 2>&1; grep -c "phase" gpurun_out/s3_c5_diag_single.log; head -12 gpurun_out/s3_c5_diag_single.log
(timeout 300 python tools/gpu_bench_cfg.py C4 128; BVHAR_NO_DIRECT_P=1 timeout 300 python tools/gpu_bench_cfg.py C4 128; timeout 300 python tools/gpu_bench_cfg.py C5 256) > gpurun_out/s3_big.log 2>&1; cat gpurun_out/s3_big.log
timeout 600 python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; tail -c 3000 gpurun_out/s3_bench.json
