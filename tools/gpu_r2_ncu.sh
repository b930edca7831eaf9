#!/bin/bash
# full ncu capture of selected kernels of one C3 sweep at 1 024 chains; summaries + hot lines (the report stays on the box)
KERN=${1:-'impact_sv|coef_chol_kernel|sv_tile_kernel|sv_solve|coef_seq_pipe'}
ncu --set full --clock-control none --import-source on -k regex:"$KERN" -s 5 -c 6 -o /tmp/r02_c3 python tools/ncu_target.py 1024 > gpurun_out/ncu_c3.log 2>&1; tail -2 gpurun_out/ncu_c3.log
python tools/ncu_summary.py /tmp/r02_c3.ncu-rep > gpurun_out/r02_ncu_kernels_c3_sel.md 2>/dev/null
for kpat in $(echo $KERN | tr '|' ' '); do python tools/ncu_hot.py /tmp/r02_c3.ncu-rep $kpat 16; done > gpurun_out/r02_ncu_hot_lines_c3.txt 2>/dev/null
ncu -i /tmp/r02_c3.ncu-rep --page details --csv 2>/dev/null | grep -E "Stall|stall|No Eligible|Eligible Warps|Issued Warp|Theoretical Occupancy|Achieved Occupancy|L1/TEX Hit|Shared.*Bank|Bank Conflicts" | head -150 > gpurun_out/r02_ncu_details_c3.txt
cat gpurun_out/r02_ncu_hot_lines_c3.txt
