#!/bin/bash
# ncu --set full of (a) the C3 contemporaneous / sequential kernels, (b) the C5 large-design kernels; summaries come back
D=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'impact_sv|coef_seq_pipe' -s 2 -c 2 -o /tmp/s3_c3 python tools/ncu_target.py 1024 > $D/s3_ncu_c3.log 2>&1; tail -1 $D/s3_ncu_c3.log
python tools/ncu_summary.py /tmp/s3_c3.ncu-rep > $D/s3_ncu_kernels_c3.md
for kp in impact_sv coef_seq_pipe; do python tools/ncu_hot.py /tmp/s3_c3.ncu-rep $kp 22; done > $D/s3_ncu_hot_c3.txt
ncu -i /tmp/s3_c3.ncu-rep --page details --csv | grep -E "Stall|No Eligible|Eligible Warps|Issued Warp|Occupancy|Bank|L1/TEX Hit|Registers|Shared Memory Config|Block Limit" > $D/s3_ncu_details_c3.txt
ncu --set full --clock-control none --import-source on -k regex:'coef_chol_cta|coef_seq_big|impact_ldlt' -s 0 -c 3 -o /tmp/s3_c5 python tools/ncu_cfg.py C5 256 1 > $D/s3_ncu_c5.log 2>&1; tail -1 $D/s3_ncu_c5.log
python tools/ncu_summary.py /tmp/s3_c5.ncu-rep > $D/s3_ncu_kernels_c5.md
for kp in coef_chol_cta coef_seq_big impact_ldlt; do python tools/ncu_hot.py /tmp/s3_c5.ncu-rep $kp 22; done > $D/s3_ncu_hot_c5.txt
ncu -i /tmp/s3_c5.ncu-rep --page details --csv | grep -E "Stall|No Eligible|Eligible Warps|Issued Warp|Occupancy|Bank|L1/TEX Hit|Registers|Block Limit" > $D/s3_ncu_details_c5.txt
cat $D/s3_ncu_kernels_c3.md $D/s3_ncu_kernels_c5.md
