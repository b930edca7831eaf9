#!/bin/bash
# what the driver runs at round end + the launch list: tests, smoke, both bench arms, ncu launch list
set -o pipefail
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log
# full capture of one sweep's kernels: the report stays on the box (too big to travel), its summaries come back
ncu --set full --clock-control none --import-source on -c 16 -s 1 -o /tmp/r01_final2 python tools/ncu_target.py 1024 > gpurun_out/ncu_final2.log 2>&1; tail -2 gpurun_out/ncu_final2.log
python tools/ncu_summary.py /tmp/r01_final2.ncu-rep > gpurun_out/r01_ncu_kernels_final.md 2>/dev/null
for kpat in dgemm_tn coef_seq_pipe impact_sv sv_mix coef_chol_kernel sv_solve; do python tools/ncu_hot.py /tmp/r01_final2.ncu-rep $kpat 10; done > gpurun_out/r01_ncu_hot_lines_final.txt 2>/dev/null
ncu -i /tmp/r01_final2.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum > gpurun_out/r01_final2_dram.csv 2>/dev/null
