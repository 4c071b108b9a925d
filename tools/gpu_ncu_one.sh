#!/bin/bash
# Full ncu capture of the scan kernels on one shape:  bash tools/gpu_ncu_one.sh "B D L" [tag]
mkdir -p gpurun_out
shape="$1"; tag=${2:-$(echo $shape | tr ' ' '_')}
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:scan_ -s 4 -c 2 -f -o gpurun_out/prof_$tag python tools/profile_one.py $shape 4 > gpurun_out/prof_$tag.log 2>&1
echo "capture $tag rc=$?"
