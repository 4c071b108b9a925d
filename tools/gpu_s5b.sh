#!/bin/bash
# session 5: backward multi-chunk kernel iteration -- quick shapes under a timeout first (a hang must not cost the box),
# scan + fused-core parity, per-shape rows, the bench line without the side measurements
mkdir -p gpurun_out
for shape in "2 8 4096" "4 8 262144" "4 256 4096" "4 64 65536"; do
  timeout -k 5 60 python tools/profile_one.py $shape 3 > gpurun_out/quick_$(echo $shape | tr ' ' '_').log 2>&1
  rc=$?; echo "quick $shape rc=$rc"
  if [ $rc -ne 0 ]; then tail -3 gpurun_out/quick_$(echo $shape | tr ' ' '_').log; echo "abort: quick shape failed"; exit 1; fi
done
timeout -k 10 900 python -m pytest tests/test_scan_gpu.py tests/test_ss2d_gpu.py -m gpu -q -x --timeout 120 --timeout-method=thread > gpurun_out/pytest_s5b.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_s5b.log
timeout -k 10 300 python tools/shape_bench.py --what scan > gpurun_out/shape_bench_s5b.log 2>&1; echo "shape rc=$?"; grep scan_bwd gpurun_out/shape_bench_s5b.log | cut -c1-120
timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s5b.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_s5b.log | cut -c1-260
