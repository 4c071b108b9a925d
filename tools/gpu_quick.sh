#!/bin/bash
# Quick GPU check of the scan kernels: a few shapes under a short timeout, then the scan parity tests and the shape bench.
mkdir -p gpurun_out
for shape in "2 8 4096" "4 8 262144" "4 256 4096" "4 1024 256"; do
  timeout -k 5 120 python tools/profile_one.py $shape 3 > gpurun_out/quick_$(echo $shape | tr ' ' '_').log 2>&1
  echo "quick $shape rc=$?"
done
timeout -k 10 900 python -m pytest tests/test_scan_gpu.py -m gpu -q -x --timeout 120 --timeout-method=thread > gpurun_out/pytest_scan.log 2>&1
echo "pytest scan rc=$?"; tail -15 gpurun_out/pytest_scan.log
timeout -k 10 600 python tools/shape_bench.py --reps 20 --what scan > gpurun_out/shape_bench.log 2>&1
echo "shape bench rc=$?"; tail -14 gpurun_out/shape_bench.log
