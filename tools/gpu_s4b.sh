#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_harness_gpu.py -x -q --timeout 300 > gpurun_out/pytest_s4b.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_s4b.log
timeout -k 10 600 python tools/harness_profile.py --steps 2 > gpurun_out/harness_profile_block.txt 2>&1; echo "profile rc=$?"
