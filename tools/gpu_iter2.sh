#!/bin/bash
# Tuning iteration: scan parity tests, channels-per-tile sweep of the pipelined kernels, one full ncu capture.
mkdir -p gpurun_out
for shape in "2 8 4096" "4 8 262144" "4 256 4096"; do
  timeout -k 5 120 python tools/profile_one.py $shape 3 > gpurun_out/quick_$(echo $shape | tr ' ' '_').log 2>&1
  echo "quick $shape rc=$?"
done
timeout -k 10 900 python -m pytest tests/test_scan_gpu.py -m gpu -q -x --timeout 120 --timeout-method=thread > gpurun_out/pytest_scan.log 2>&1
echo "pytest scan rc=$?"; tail -5 gpurun_out/pytest_scan.log
bash tools/gpu_sweep.sh "$@"
bash tools/gpu_ncu_one.sh "4 64 65536" pipe_4_64_65536
