#!/bin/bash
# ncu launch list of the bench command (every launch: device time + DRAM bytes) and full captures of the scan kernels
# on three shapes.  Run under gpurun.  Keeps gpurun_out/ under the 64 MiB merge limit.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-core --no-stft --no-cpu-baseline"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --graph-profiling node -c 4000 --csv --log-file gpurun_out/launches_raw.csv $BENCH > gpurun_out/launches_run.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_raw.csv)"
if ! grep -q scan_ gpurun_out/launches_raw.csv; then
  # graph nodes not visible to this ncu: same command with the step launched eagerly
  timeout -k 10 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -c 4000 --csv --log-file gpurun_out/launches_raw.csv $BENCH --no-graph > gpurun_out/launches_run.log 2>&1
  echo "launch list (eager) rc=$? lines=$(wc -l < gpurun_out/launches_raw.csv)"
fi
python tools/launch_list.py gpurun_out/launches_raw.csv gpurun_out/launches.csv gpurun_out/dominant_kernel_traffic.json
rm -f gpurun_out/launches_raw.csv
for shape in "4 8 262144" "4 64 65536" "4 256 4096"; do
  tag=$(echo $shape | tr ' ' '_')
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:scan_ -s 4 -c 2 -f -o gpurun_out/prof_$tag python tools/profile_one.py $shape 4 > gpurun_out/prof_$tag.log 2>&1
  echo "capture $tag rc=$?"
done
du -sh gpurun_out
