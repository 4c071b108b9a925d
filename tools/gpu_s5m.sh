#!/bin/bash
# session 5: programmatic dependent launch ON in the product library -- bench line x3 (and the tuning build with VMASR_PDL=0 for
# the same box), then the whole GPU suite and the sanitizer's multi-tile / fused shapes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2 3; do
timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s5m_$i.log 2>&1; echo "bench (product, PDL on) rc=$?"; tail -1 gpurun_out/bench_s5m_$i.log | cut -c1-200
done
for cfg in "VMASR_PDL=0" "VMASR_PDL=1" "VMASR_PDL=0" "VMASR_PDL=1"; do
  env $cfg VMASR_B200_LIBRARY=vm_asr_b200/lib_tuning/libvmasr_b200.so timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s5m_t.log 2>&1
  echo "== tuning build $cfg: bench $(tail -1 gpurun_out/bench_s5m_t.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["frac_of_hbm_peak"])' 2>&1 | tail -1)"
done
timeout -k 10 1800 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for tool in memcheck racecheck; do
  for shape in "4 64 16384" "1 8 4112"; do
    tag=$(echo $shape | tr ' ' '_')
    timeout -k 10 400 compute-sanitizer --tool $tool --error-exitcode 9 python tools/profile_one.py $shape 2 > gpurun_out/sanitize_${tool}_$tag.log 2>&1
    echo "$tool scan $shape rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_$tag.log | tail -1)"
  done
  timeout -k 10 400 compute-sanitizer --tool $tool --error-exitcode 9 python tools/profile_fused.py 1 4 48 64 > gpurun_out/sanitize_${tool}_fused_1_4_48_64.log 2>&1
  echo "$tool fused+stft 1 4 48 64 rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_fused_1_4_48_64.log | tail -1)"
done
