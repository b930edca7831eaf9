#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/s3_bench_mid.json 2> gpurun_out/s3_bench_mid.err; tail -3 gpurun_out/s3_bench_mid.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/s3_bench_mid.json") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_all"], "roof", d["roofline"]["frac"], "launches", d["gpu_launches"])
print("roll", d["forecast_roll"]["value"], d["forecast_roll"]["seconds"])
print(d["roofline"]["kernel_ms_per_step"])
PY
