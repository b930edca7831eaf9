#!/bin/bash
# round 2: both bench arms on one GPU
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
tail -c 400 gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench.json") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_all"], "roofline", d["roofline"]["frac"], d["roofline"]["sweep_frac_of_peak"])
print("roll", {k: v for k, v in d["forecast_roll"].items() if k not in ("roofline", "config")})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["value_efficient_formulation"])
print(d["clocks"], d["roofline"]["fp64_peaks"])
PY
