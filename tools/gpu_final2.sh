#!/bin/bash
# Round-end consolidation (session 5 tree): the whole GPU suite, smoke, the bench line, the reference arm, per-shape rows,
# sanitizer over every fast path, ncu launch list of the bench command, full ncu captures of the scan kernels on three shapes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout -k 10 1800 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -k 10 900 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_final.log | cut -c1-400
timeout -k 10 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference_final.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_reference_final.log | cut -c1-300
timeout -k 10 300 python tools/shape_bench.py > gpurun_out/shape_bench_final.log 2>&1; echo "shape rc=$?"
tail -3 gpurun_out/vs_reference_cuda.jsonl | cut -c1-200   # (written by tests/test_vs_reference_cuda_gpu.py)
bash tools/gpu_sanitize.sh
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-core --no-stft --no-cpu-baseline"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --graph-profiling node -c 4000 --csv --log-file gpurun_out/launches_raw.csv $BENCH > gpurun_out/launches_run.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_raw.csv)"
python tools/launch_list.py gpurun_out/launches_raw.csv gpurun_out/launches.csv gpurun_out/dominant_kernel_traffic.json
rm -f gpurun_out/launches_raw.csv
for shape in "4 64 65536" "4 256 4096" "4 8 262144"; do
  tag=$(echo $shape | tr ' ' '_')
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:scan_ -s 4 -c 2 -f -o gpurun_out/prof_$tag python tools/profile_one.py $shape 4 > gpurun_out/prof_$tag.log 2>&1
  echo "capture $tag rc=$?"
done
du -sh gpurun_out
