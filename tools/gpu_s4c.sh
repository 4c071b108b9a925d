#!/bin/bash
# ncu full captures of the head / tail kernels on two config shapes
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
for shape in "4 32 128 128" "4 2 512 512" "4 128 32 32"; do
  tag=$(echo $shape | tr ' ' '_')
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"dwconv|outnorm" -c 8 -f -o gpurun_out/headtail_$tag python tools/profile_fused.py $shape > gpurun_out/headtail_$tag.log 2>&1
  echo "capture $tag rc=$?"
done
du -sh gpurun_out
