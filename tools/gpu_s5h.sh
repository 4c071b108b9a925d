#!/bin/bash
# session 5: forward CTAs that take several tiles -- parity, bench line, per-shape rows, multi-tile vs one-tile (tuning build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for shape in "2 8 4096" "4 8 262144" "4 256 4096" "4 64 65536"; do
  timeout -k 5 60 python tools/profile_one.py $shape 3 > gpurun_out/quick_$(echo $shape | tr ' ' '_').log 2>&1
  rc=$?; echo "quick $shape rc=$rc"
  if [ $rc -ne 0 ]; then tail -3 gpurun_out/quick_$(echo $shape | tr ' ' '_').log; echo "abort: quick shape failed"; exit 1; fi
done
timeout -k 10 900 python -m pytest tests/test_scan_gpu.py tests/test_ss2d_gpu.py -m gpu -q -x --timeout 120 --timeout-method=thread > gpurun_out/pytest_s5h.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_s5h.log
timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s5h.log 2>&1; echo "bench (product) rc=$?"; tail -1 gpurun_out/bench_s5h.log | cut -c1-200
timeout -k 10 300 python tools/shape_bench.py --what scan > gpurun_out/shape_bench_s5h.log 2>&1; grep "scan_fwd" gpurun_out/shape_bench_s5h.log | cut -c1-120
export VMASR_B200_LIBRARY=vm_asr_b200/lib_tuning/libvmasr_b200.so
i=0
for cfg in "X=0" "VMASR_FWD_PERSIST_ROUNDS=0" "VMASR_FWD_PERSIST_ROUNDS=6"; do
  i=$((i+1))
  env $cfg timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s5h_$i.log 2>&1
  echo "== $cfg: bench $(tail -1 gpurun_out/bench_s5h_$i.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["frac_of_hbm_peak"])' 2>&1 | tail -1)"
done
