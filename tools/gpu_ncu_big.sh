#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:'coef_chol_smem|coef_seq_time' -c 2 -o gpurun_out/r01_c4_big python tools/ncu_cfg.py C4 128 1 > gpurun_out/ncu_c4.log 2>&1; tail -2 gpurun_out/ncu_c4.log
BVHAR_CHOL_GLOBAL=1 ncu --set full --clock-control none --import-source on -k regex:'coef_chol_cta' -c 1 -o gpurun_out/r01_c4_cholg python tools/ncu_cfg.py C4 128 1 > gpurun_out/ncu_c4g.log 2>&1; tail -2 gpurun_out/ncu_c4g.log
