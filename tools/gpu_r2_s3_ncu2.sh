#!/bin/bash
D=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'coef_chol_cta' -s 0 -c 1 -o /tmp/s3_chol python tools/ncu_cfg.py ${1:-C5} ${2:-256} 1 > $D/s3_ncu_chol.log 2>&1; tail -1 $D/s3_ncu_chol.log
python tools/ncu_summary.py /tmp/s3_chol.ncu-rep > $D/s3_ncu_kernels_chol.md
python tools/ncu_hot.py /tmp/s3_chol.ncu-rep coef_chol_cta 30 > $D/s3_ncu_hot_chol.txt
ncu -i /tmp/s3_chol.ncu-rep --page details --csv | grep -E "Stall|No Eligible|Eligible Warps|Issued Warp|Occupancy|Bank|L1/TEX Hit|Registers|Block Limit|Wavefront|Pipe|Throughput" > $D/s3_ncu_details_chol.txt
cat $D/s3_ncu_kernels_chol.md; cat $D/s3_ncu_hot_chol.txt
