#!/bin/bash
# compute-sanitizer passes over small instances of the configurations (memcheck, racecheck, initcheck)
CS=/usr/local/cuda/bin/compute-sanitizer
for cfg in ${CFGS:-"C5 8 2" "C4 8 2" "C3 8 2" "C1 4 2"}; do
  for tool in ${TOOLS:-memcheck racecheck}; do
    echo "=== compute-sanitizer --tool $tool : gpu_bench_cfg.py $cfg"
    timeout 1200 $CS --tool $tool --print-limit 8 python tools/gpu_bench_cfg.py $cfg 2>&1 | grep -vE "^   [a-z_]+ +[0-9.]+ ms/sweep" | tail -25
  done
done
