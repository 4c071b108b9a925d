#!/bin/bash
# Round 2, session n: phase timelines after the epoch change; PDL on the bench step (tuning build).
mkdir -p gpurun_out
export VMASR_B200_LIBRARY=$PWD/vm_asr_b200/lib_tuning/libvmasr_b200.so
rm -f gpurun_out/timeline_n.log
for shape in "4 64 65536" "4 256 4096" "4 8 262144"; do
  timeout -k 10 120 python tools/timeline.py $shape 2>&1 | tee -a gpurun_out/timeline_n.log
done
VMASR_PDL=1 timeout -k 10 600 python bench.py --steps 30 > gpurun_out/bench_pdl.log 2>&1; echo "bench pdl rc=$?"; tail -1 gpurun_out/bench_pdl.log | cut -c1-260
