#!/bin/bash
# compute-sanitizer over the scan kernels on small shapes of every fast path, the fused SS2D core (time-reversed directions,
# load-add-store second pass, map kernels) and the STFT kernels (memcheck, then racecheck on shared memory).
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  # (the third shape has more tiles than resident CTAs: the persistent backward's CTAs walk two tiles each)
  for shape in "1 8 4112" "2 16 8192" "4 64 16384" "1 8 1024" "1 32 256"; do
    tag=$(echo $shape | tr ' ' '_')
    timeout -k 10 300 compute-sanitizer --tool $tool --error-exitcode 9 python tools/profile_one.py $shape 1 > gpurun_out/sanitize_${tool}_$tag.log 2>&1
    echo "$tool scan $shape rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_$tag.log | tail -1)"
  done
  for shape in "1 4 48 64" "2 8 16 16"; do
    tag=$(echo $shape | tr ' ' '_')
    timeout -k 10 400 compute-sanitizer --tool $tool --error-exitcode 9 python tools/profile_fused.py $shape > gpurun_out/sanitize_${tool}_fused_$tag.log 2>&1
    echo "$tool fused+stft $shape rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_fused_$tag.log | tail -1)"
  done
done
