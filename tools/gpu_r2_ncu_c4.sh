#!/bin/bash
# round 2: full ncu capture of the large-design (C4, forecast_roll shape: 128 units) sweep kernels; summaries come back
ncu --set full --clock-control none --import-source on -k regex:'coef_chol_cta|coef_seq_big|invert_diag|sv_wmix|impact_sv|impact_rows|dgemm_tn|prior_kernel' -s 10 -c 12 -o /tmp/r02_c4 python tools/ncu_cfg.py C4 128 2 > gpurun_out/ncu_c4_r02.log 2>&1; tail -2 gpurun_out/ncu_c4_r02.log
python tools/ncu_summary.py /tmp/r02_c4.ncu-rep > gpurun_out/r02_ncu_kernels_c4_128units.md 2>/dev/null
for kpat in coef_chol_cta2 coef_seq_big; do python tools/ncu_hot.py /tmp/r02_c4.ncu-rep $kpat 14; done > gpurun_out/r02_ncu_hot_lines_c4.txt 2>/dev/null
ncu -i /tmp/r02_c4.ncu-rep --page details --csv 2>/dev/null | grep -E "coef_chol_cta2|coef_seq_big" | grep -E "Stall|Warp Cycles Per|Eligible|Issued Warp|No Eligible|Registers|Shared Memory Config|Theoretical Occ|Achieved Occ|Block Limit" | cut -c1-400 > gpurun_out/r02_ncu_details_c4.txt
cat gpurun_out/r02_ncu_kernels_c4_128units.md
head -40 gpurun_out/r02_ncu_hot_lines_c4.txt
