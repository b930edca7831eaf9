#!/bin/bash
# last state of the round: the driver's own sequence (GPU tests, smoke, both bench arms)
set -o pipefail
D=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 300 python bench.py --impl reference > $D/r02_last_ref.json 2> $D/r02_last_ref.err; tail -c 300 $D/r02_last_ref.json
timeout 400 python bench.py > $D/r02_last.json 2> $D/r02_last.err; tail -3 $D/r02_last.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_last.json") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["frac"], "launches", d["gpu_launches"], "roll", d["forecast_roll"]["value"], "clocks", d["clocks"])
PY
