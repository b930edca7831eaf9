#!/bin/bash
# round 2: the full-size parity tests verbosely, the select forecasters, then the whole gpu suite
python -m pytest tests/test_gpu_baseline_parity.py -m gpu -q -s 2>&1 | grep -E "cond|FAILED|Error|passed|failed" > gpurun_out/r2_parity.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "select" 2>&1 | tail -15 > gpurun_out/r2_select.log
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_gputests.log
cat gpurun_out/r2_parity.log gpurun_out/r2_select.log gpurun_out/r2_gputests.log
