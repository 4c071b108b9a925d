#!/bin/bash
# session 6: dB / dC buffers cleared by the forward launches (vmasr_scan_params.zero_ptr): tests, same-box A/B of the bench step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_scan_gpu.py -x -q -m gpu --timeout 300 > gpurun_out/pytest_s6b.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_s6b.log
for i in 1 2; do
  timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s6b_fwdzero_$i.log 2>&1; echo "bench fwd-zero rc=$?"; tail -1 gpurun_out/bench_s6b_fwdzero_$i.log | cut -c1-160
  VMASR_BENCH_NO_FWD_ZERO=1 timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s6b_memset_$i.log 2>&1; echo "bench memset rc=$?"; tail -1 gpurun_out/bench_s6b_memset_$i.log | cut -c1-160
done
