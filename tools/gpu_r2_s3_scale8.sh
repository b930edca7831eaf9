#!/bin/bash
# N = 8 (and 4) runs of bench.py as the driver launches it: weak value, strong block, e2e, forecast_roll
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s3_scale_$n.json 2> gpurun_out/s3_scale_$n.err
  python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/s3_scale_$n.json") if x.startswith("{")][-1])
print($n, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("seconds_all"))
print("   strong", d["strong"])
print("   roll", {k: v for k, v in d["forecast_roll"].items() if k in ("value", "seconds", "windows_total")})
PY
done
