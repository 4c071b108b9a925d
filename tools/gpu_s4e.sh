#!/bin/bash
# ncu full capture of one kernel family on three config shapes:  bash tools/gpu_s4e.sh <kernel regex> <tag>
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
for shape in "4 32 128 128" "4 2 512 512" "4 128 32 32"; do
  tag=$(echo $shape | tr ' ' '_')
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"$1" -c ${3:-2} -f -o gpurun_out/$2_$tag python tools/profile_fused.py $shape > gpurun_out/$2_$tag.log 2>&1
  echo "capture $tag rc=$?"
done
