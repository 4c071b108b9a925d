#!/bin/bash
# session 6: griddepcontrol.wait behind the CTA's set-up (no tensor-map prefetch any more) -- same-box A/B on the tuning build
# (VMASR_PDL_X=1: wait at the top as before), product library, quick parity pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_scan_gpu.py tests/test_ss2d_gpu.py -x -q -m gpu --timeout 300 -k "grouped or golden or zero or dbdc or ragged or reverse or accumulate or core or graph" > gpurun_out/pytest_s6e.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_s6e.log
for round in 1 2 3; do
for x in 0 1; do
  VMASR_PDL_X=$x VMASR_B200_LIBRARY=vm_asr_b200/lib_tuning/libvmasr_b200.so timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s6e_t.log 2>&1
  echo "== tuning build VMASR_PDL_X=$x: $(tail -1 gpurun_out/bench_s6e_t.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["frac_of_hbm_peak"])' 2>&1 | tail -1)"
done
done
for i in 1 2; do
timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s6e_$i.log 2>&1; echo "bench (product) rc=$?"; tail -1 gpurun_out/bench_s6e_$i.log | cut -c1-170
done
