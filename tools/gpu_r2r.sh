#!/bin/bash
mkdir -p gpurun_out
TRIALS=4 timeout -k 10 600 python tools/debug_harness2.py > gpurun_out/debug_harness2.log 2>&1; echo "dbg rc=$?"; cat gpurun_out/debug_harness2.log | tail -60
for i in 1 2 3; do timeout -k 10 300 python -m pytest tests/test_harness_gpu.py -q -x --timeout 300 2>&1 | tail -3; done
timeout -k 10 900 python -m pytest tests/test_ss2d_gpu.py -x -q -k "merge_norm or core_out" --timeout 300 > gpurun_out/pytest_tail.log 2>&1; echo "tail rc=$?"; tail -8 gpurun_out/pytest_tail.log
