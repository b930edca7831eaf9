#!/bin/bash
# full ncu capture of the large-design (C4) kernels; the report stays on the box, its summaries come back
ncu --set full --clock-control none --import-source on -k regex:'coef_chol_cta|invert_diag|coef_seq_time|coef_pbuild|impact_sv|dgemm_tn' -s 1 -c 8 -o /tmp/r01_c4_final python tools/ncu_cfg.py C4 512 1 > gpurun_out/ncu_c4.log 2>&1; tail -2 gpurun_out/ncu_c4.log
python tools/ncu_summary.py /tmp/r01_c4_final.ncu-rep > gpurun_out/r01_ncu_kernels_c4_final.md 2>/dev/null
for kpat in coef_chol_cta coef_seq_time; do python tools/ncu_hot.py /tmp/r01_c4_final.ncu-rep $kpat 12; done > gpurun_out/r01_ncu_hot_lines_c4_final.txt 2>/dev/null
python tools/gpu_bench_cfg.py C4 1024 3 > gpurun_out/q4_final.log 2>&1; python tools/gpu_bench_cfg.py C5 256 3 > gpurun_out/q5_final.log 2>&1; python tools/gpu_bench_cfg.py C1 1024 10 > gpurun_out/q1_final.log 2>&1; python tools/gpu_bench_cfg.py C2 64 10 > gpurun_out/q2_final.log 2>&1
cat gpurun_out/q4_final.log | head -8
