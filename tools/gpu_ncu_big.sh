#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:'dgemm_tn' -s 2 -c 1 -o gpurun_out/r01_c4_gram python tools/ncu_cfg.py C4 512 1 > gpurun_out/ncu_c4.log 2>&1; tail -2 gpurun_out/ncu_c4.log
