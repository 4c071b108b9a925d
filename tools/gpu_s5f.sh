#!/bin/bash
# session 5: programmatic dependent launch with the persistent backward; timelines of the persistent backward
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export VMASR_B200_LIBRARY=vm_asr_b200/lib_tuning/libvmasr_b200.so
i=0
for cfg in "X=0" "VMASR_PDL=1"; do
  i=$((i+1))
  env $cfg timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s5f_$i.log 2>&1
  echo "== $cfg: bench $(tail -1 gpurun_out/bench_s5f_$i.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["frac_of_hbm_peak"])' 2>&1 | tail -1)"
done
rm -f gpurun_out/timeline_s5f.txt
for shape in "4 64 65536" "4 128 16384" "4 256 4096"; do
  timeout -k 5 120 python tools/timeline.py $shape >> gpurun_out/timeline_s5f.txt 2>&1
done
grep -A18 "== bwd" gpurun_out/timeline_s5f.txt | cut -c1-150
