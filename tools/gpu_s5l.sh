#!/bin/bash
# session 5: the step's zero-fill on a side branch of the graph (bench.py) -- bench line, product library; tuning build with and without PDL
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2; do
timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s5l_$i.log 2>&1; echo "bench (product) rc=$?"; tail -1 gpurun_out/bench_s5l_$i.log | cut -c1-200
done
export VMASR_B200_LIBRARY=vm_asr_b200/lib_tuning/libvmasr_b200.so
for cfg in "X=0" "VMASR_PDL=1"; do
  env $cfg timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s5l_t.log 2>&1
  echo "== $cfg: bench $(tail -1 gpurun_out/bench_s5l_t.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["frac_of_hbm_peak"])' 2>&1 | tail -1)"
done
