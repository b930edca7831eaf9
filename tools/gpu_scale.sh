#!/bin/bash
# weak-scaling run of bench.py at N = 1, 2, 4, 8 as the driver launches it
N=${1:-8}
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; fi
  python - <<PY
import json
l=[x for x in open("gpurun_out/scale_$n.json") if x.startswith("{")]
d=json.loads(l[-1]); print($n, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("clocks"))
PY
done
