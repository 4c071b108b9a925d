#!/bin/bash
# Tuning iteration with an environment prefix for the FIRST phase (quick shapes + parity tests), then a sweep:
#   bash tools/gpu_iter3.sh "ENV=1 ENV2=x" "sweep cfg 1" "sweep cfg 2" ...
mkdir -p gpurun_out
pre="$1"; shift
for shape in "2 8 4096" "4 8 262144" "4 256 4096" "4 64 65536"; do
  env $pre timeout -k 5 60 python tools/profile_one.py $shape 3 > gpurun_out/quick_$(echo $shape | tr ' ' '_').log 2>&1
  rc=$?; echo "quick $shape rc=$rc"
  if [ $rc -ne 0 ]; then tail -3 gpurun_out/quick_$(echo $shape | tr ' ' '_').log; echo "abort: quick shape failed"; exit 1; fi
done
env $pre timeout -k 10 900 python -m pytest tests/test_scan_gpu.py -m gpu -q -x --timeout 120 --timeout-method=thread > gpurun_out/pytest_scan.log 2>&1
echo "pytest scan ($pre) rc=$?"; tail -5 gpurun_out/pytest_scan.log
bash tools/gpu_sweep.sh "$@"
