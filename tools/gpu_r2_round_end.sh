#!/bin/bash
# round-2 final state: what the driver runs at round end + the evidence for profiles/ (launch list, full ncu capture, per-config timings)
set -o pipefail
D=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --impl reference > $D/r02_bench_ref.json 2> $D/r02_bench_ref.err; tail -c 400 $D/r02_bench_ref.json
python bench.py > $D/r02_bench.json 2> $D/r02_bench.err; tail -3 $D/r02_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench.json") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_all"], "roof", d["roofline"]["frac"], "launches", d["gpu_launches"], "roll", d["forecast_roll"]["value"], "clocks", d["clocks"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $D/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roll --no-strong > $D/ncu_bench.log 2>&1; tail -1 $D/ncu_bench.log
# full capture of one sweep's kernels (eager launches of the 2nd sweep): the report stays on the box, its summaries come back
ncu --set full --clock-control none --import-source on -c 20 -s 20 -o /tmp/r02_final python tools/ncu_target.py 1024 > $D/ncu_final.log 2>&1; tail -1 $D/ncu_final.log
python tools/ncu_summary.py /tmp/r02_final.ncu-rep > $D/r02_ncu_kernels_final.md 2>/dev/null
for kpat in dgemm_tn_bulk coef_seq_pipe impact_sv impact_rows sv_tile coef_chol_kernel sv_solve prior_kernel coef_pbuild; do python tools/ncu_hot.py /tmp/r02_final.ncu-rep $kpat 10; done > $D/r02_ncu_hot_lines_final.txt 2>/dev/null
ncu -i /tmp/r02_final.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum > $D/r02_final_dram.csv 2>/dev/null
for c in "C1 1024 10" "C2 64 10" "C4 128" "C4 1024 3" "C5 256 3"; do python tools/gpu_bench_cfg.py $c; done > $D/r02_configs.log 2>&1
cat $D/r02_ncu_kernels_final.md
