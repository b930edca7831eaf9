#!/bin/bash
D=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'chol_panel_kernel' -s 0 -c 1 -o /tmp/s3_cp python tools/ncu_cfg.py C4 128 1 > $D/s3_ncu_cp.log 2>&1; tail -1 $D/s3_ncu_cp.log
python tools/ncu_summary.py /tmp/s3_cp.ncu-rep
python tools/ncu_hot.py /tmp/s3_cp.ncu-rep chol_panel 24
ncu -i /tmp/s3_cp.ncu-rep --page details --csv | grep -E "No Eligible|Eligible Warps|Issued Warp|Block Limit|Achieved Occupancy|L1/TEX Hit|Local" | cut -d, -f12-14
