#!/bin/bash
# N = 8 run of bench.py as the driver launches it, after the set-up work: weak value, strong block, e2e, forecast_roll
n=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_scale_8_last.json 2> gpurun_out/r02_scale_8_last.err
python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/r02_scale_8_last.json") if x.startswith("{")][-1])
print($n, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("seconds_all"))
print("   strong", d["strong"])
print("   roll", {k: v for k, v in d["forecast_roll"].items() if k in ("value", "seconds", "windows_total")})
PY
