#!/bin/bash
mkdir -p gpurun_out
for dt in float32 float16; do timeout -k 10 300 python tools/tail_bench.py --dtype $dt 2>&1 | tee -a gpurun_out/tail_bench.jsonl | cut -c1-420; done
for tool in memcheck racecheck; do
  for shape in "1 4 48 64" "2 8 16 16"; do
    tag=$(echo $shape | tr ' ' '_')
    timeout -k 10 400 compute-sanitizer --tool $tool --error-exitcode 9 python tools/profile_fused.py $shape > gpurun_out/sanitize_${tool}_fused_$tag.log 2>&1
    echo "$tool fused+tail+stft $shape rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_fused_$tag.log | tail -1)"
  done
done
