#!/bin/bash
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err
tail -c 300 gpurun_out/r2_scale_$N.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_scale_$N.json") if l.startswith("{")][-1])
print("N", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["seconds_all"])
print("strong", d["strong"])
print("roll", {k: v for k, v in d["forecast_roll"].items() if k not in ("roofline", "config", "cpu_baseline")})
PY
