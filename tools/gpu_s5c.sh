#!/bin/bash
# session 5: persistent backward with tickets -- quick shapes under a timeout, scan parity, per-shape rows, bench line, timelines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for shape in "2 8 4096" "4 8 262144" "4 256 4096" "4 64 65536"; do
  timeout -k 5 60 python tools/profile_one.py $shape 3 > gpurun_out/quick_$(echo $shape | tr ' ' '_').log 2>&1
  rc=$?; echo "quick $shape rc=$rc"
  if [ $rc -ne 0 ]; then tail -3 gpurun_out/quick_$(echo $shape | tr ' ' '_').log; echo "abort: quick shape failed"; exit 1; fi
done
timeout -k 10 600 python -m pytest ${PYTEST_FILES:-tests/test_scan_gpu.py} -m gpu -q -x --timeout 120 --timeout-method=thread > gpurun_out/pytest_s5c.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_s5c.log
timeout -k 10 300 python tools/shape_bench.py --what scan > gpurun_out/shape_bench_s5c.log 2>&1; echo "shape rc=$?"; grep "scan_" gpurun_out/shape_bench_s5c.log | cut -c1-120
timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s5c.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_s5c.log | cut -c1-260
rm -f gpurun_out/timeline_s5c.txt
for shape in "4 64 65536" "4 8 262144" "4 128 16384"; do
  VMASR_B200_LIBRARY=vm_asr_b200/lib_tuning/libvmasr_b200.so timeout -k 5 120 python tools/timeline.py $shape >> gpurun_out/timeline_s5c.txt 2>&1
done
grep -A16 "== bwd" gpurun_out/timeline_s5c.txt | head -60
