#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> '<command>'   - retries while the pod answers busy (exit 3 / transient)
T=$1; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gpurun: still busy after 12 tries"; exit 3
