#!/bin/bash
# what the driver runs at round end + the launch list: tests, smoke, both bench arms, ncu launch list
set -o pipefail
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log
