#!/bin/bash
# session 6, round-end consolidation with the final library: whole GPU suite, smoke, bench line, reference arm, pair launches,
# the two B = 8 configs, sanitizer over the scan fast paths, ncu launch list of the bench command, full ncu captures on two shapes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/prof_4_64_65536.ncu-rep gpurun_out/prof_4_8_262144.ncu-rep
timeout -k 10 1500 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -k 10 900 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_final.log | cut -c1-400
timeout -k 10 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference_final.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_reference_final.log | cut -c1-300
timeout -k 10 300 python tools/shape_bench.py --what scan --pairs > gpurun_out/shape_pairs_final.log 2>&1; echo "pairs rc=$?"; grep scan_ gpurun_out/shape_pairs_final.log | cut -c1-130
timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline --workload vm_asr_48k_16k_nfft2048 > gpurun_out/bench_nfft2048.log 2>&1; echo "nfft2048 rc=$?"; tail -1 gpurun_out/bench_nfft2048.log | cut -c1-200
timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline --workload vm_asr_48k_16k_MPD_VSSM32 > gpurun_out/bench_vssm32.log 2>&1; echo "vssm32 rc=$?"; tail -1 gpurun_out/bench_vssm32.log | cut -c1-200
for tool in memcheck racecheck; do
  for shape in "1 8 4112" "4 64 16384" "1 8 1024"; do
    tag=$(echo $shape | tr ' ' '_')
    timeout -k 10 300 compute-sanitizer --tool $tool --error-exitcode 9 python tools/profile_one.py $shape 1 > gpurun_out/sanitize_${tool}_$tag.log 2>&1
    echo "$tool scan $shape rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_$tag.log | tail -1)"
  done
done
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-core --no-stft --no-cpu-baseline"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --graph-profiling node -c 4000 --csv --log-file gpurun_out/launches_raw.csv $BENCH > gpurun_out/launches_run.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_raw.csv)"
python tools/launch_list.py gpurun_out/launches_raw.csv gpurun_out/launches.csv gpurun_out/dominant_kernel_traffic.json
rm -f gpurun_out/launches_raw.csv
for shape in "4 64 65536" "4 8 262144"; do
  tag=$(echo $shape | tr ' ' '_')
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:scan_ -s 4 -c 2 -f -o gpurun_out/prof_$tag python tools/profile_one.py $shape 4 > gpurun_out/prof_$tag.log 2>&1
  echo "capture $tag rc=$?"
done
