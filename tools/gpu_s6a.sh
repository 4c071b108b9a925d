#!/bin/bash
# session 6: dB / dC stored instead of zero-filled + accumulated on the C = 2 maps (VMASR_SCAN_DBDC_STORE): tests, same-box A/B of the bench step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_scan_gpu.py -x -q -m gpu --timeout 300 -k "dbdc or grouped or accumulate or ragged or golden or prepared or autograd" > gpurun_out/pytest_s6a.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_s6a.log
for i in 1 2; do
  timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s6a_store_$i.log 2>&1; echo "bench store rc=$?"; tail -1 gpurun_out/bench_s6a_store_$i.log | cut -c1-160
  VMASR_BENCH_NO_DBDC_STORE=1 timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s6a_nostore_$i.log 2>&1; echo "bench no-store rc=$?"; tail -1 gpurun_out/bench_s6a_nostore_$i.log | cut -c1-160
done
