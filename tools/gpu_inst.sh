#!/bin/bash
# Dynamic instruction counts of the scan kernels on a few shapes (metrics-only ncu pass), then the shape bench.
mkdir -p gpurun_out
: > gpurun_out/inst.csv
for shape in "4 8 262144" "4 64 65536" "4 256 4096" "4 512 1024"; do
  timeout -k 5 200 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:scan_ -s 4 -c 2 --csv python tools/profile_one.py $shape 4 2>/dev/null | grep -E "scan_" | awk -F'","' -v s="$shape" '{print s "," $5 "," $(NF-2) "," $(NF)}' >> gpurun_out/inst.csv
done
cat gpurun_out/inst.csv
timeout -k 10 600 python tools/shape_bench.py --reps 20 --what scan > gpurun_out/shape_bench.log 2>&1
echo "shape bench rc=$?"; tail -14 gpurun_out/shape_bench.log
