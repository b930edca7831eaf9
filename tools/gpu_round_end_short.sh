#!/bin/bash
# tests, smoke, both bench arms (no ncu)
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["seconds_all"], d["roofline"]["frac"], d["forecast_roll"]["value"], d["clocks"])
PY
