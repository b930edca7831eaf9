#!/bin/bash
# full ncu capture of the persistent bulk-copy contraction (G2, G1, G3 launches of two C3 sweeps at 1 024 chains)
ncu --set full --clock-control none --import-source on -k regex:"dgemm_tn_bulk" -s 1 -c 6 -o /tmp/r02_g1 python tools/ncu_target.py 1024 > gpurun_out/ncu_g1.log 2>&1; tail -2 gpurun_out/ncu_g1.log
python tools/ncu_summary.py /tmp/r02_g1.ncu-rep > gpurun_out/r02_ncu_kernels_g1.md 2>/dev/null
python tools/ncu_hot.py /tmp/r02_g1.ncu-rep dgemm_tn_bulk 14 > gpurun_out/r02_ncu_hot_lines_g1.txt 2>/dev/null
ncu -i /tmp/r02_g1.ncu-rep --page details --csv 2>/dev/null | grep -E "Stall|No Eligible|Eligible Warps|Issued Warp|Achieved Occupancy|Bank Conflicts|Pipe|L2 Hit|DRAM Throughput" | head -80 > gpurun_out/r02_ncu_details_g1.txt
cat gpurun_out/r02_ncu_kernels_g1.md
