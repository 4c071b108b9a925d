#!/bin/bash
# Round 2, eighth GPU session: sanitizer, ncu launch list of the bench command, full captures of the scan kernels (for profiles/).
mkdir -p gpurun_out
bash tools/gpu_sanitize.sh 2>&1 | tee gpurun_out/sanitizer_summary.txt
rm -f gpurun_out/*.ncu-rep
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-core --no-stft --no-graph"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -c 4000 --csv --log-file gpurun_out/launches_raw.csv $BENCH > gpurun_out/launches_run.log 2>&1
echo "launch list (eager) rc=$? lines=$(wc -l < gpurun_out/launches_raw.csv)"
python tools/launch_list.py gpurun_out/launches_raw.csv gpurun_out/launches.csv gpurun_out/dominant_kernel_traffic.json
tail -12 gpurun_out/launches.csv
rm -f gpurun_out/launches_raw.csv
for shape in "4 64 65536" "4 256 4096"; do
  tag=$(echo $shape | tr ' ' '_')
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:scan_ -s 4 -c 2 -f -o gpurun_out/prof_r2b_$tag python tools/profile_one.py $shape 4 > gpurun_out/prof_r2b_$tag.log 2>&1
  echo "capture $tag rc=$?"
done
timeout -k 10 600 ncu --set full --clock-control none -k regex:"stft|synth|finalize|istft" -s 5 -c 5 -f -o gpurun_out/prof_r2b_stft python tools/profile_stft.py 4 > gpurun_out/prof_r2b_stft.log 2>&1
echo "capture stft rc=$?"
du -sh gpurun_out
