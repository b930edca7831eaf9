#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:'coef_chol_cta|coef_seq_time|coef_pbuild|impact_sv' -c 4 -o gpurun_out/r01_c4_big2 python tools/ncu_cfg.py C4 512 1 > gpurun_out/ncu_c4.log 2>&1; tail -2 gpurun_out/ncu_c4.log
