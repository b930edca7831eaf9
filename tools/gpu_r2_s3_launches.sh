#!/bin/bash
# launch list (ncu time per launch) of one sweep of a configuration
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/s3_launches_$1.csv python tools/ncu_cfg.py $1 $2 1 > gpurun_out/s3_launches_$1.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/s3_launches_$1.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows:
    name=r[4].split("(")[0][-60:]; print(f"{name:62s} grid {r[7]:>18s} {float(r[-1])/1e3:10.1f} us")
PY
