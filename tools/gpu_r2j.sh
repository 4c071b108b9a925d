#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python tools/debug_harness.py > gpurun_out/debug_harness.log 2>&1
echo "rc=$?"; tail -80 gpurun_out/debug_harness.log
