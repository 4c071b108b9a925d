#!/bin/bash
# session 5, last call: final library -- scan + fused-core + harness parity, racecheck / memcheck on the multi-tile shape and the
# fused core, then the records: bench line (all legs), reference arm, per-shape rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_scan_gpu.py tests/test_ss2d_gpu.py tests/test_harness_gpu.py tests/test_model_dropin_gpu.py -m gpu -q -x --timeout 300 --timeout-method=thread > gpurun_out/pytest_s5j.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_s5j.log
for tool in memcheck racecheck; do
  for shape in "4 64 16384" "2 16 8192"; do
    tag=$(echo $shape | tr ' ' '_')
    timeout -k 10 400 compute-sanitizer --tool $tool --error-exitcode 9 python tools/profile_one.py $shape 1 > gpurun_out/sanitize_${tool}_$tag.log 2>&1
    echo "$tool scan $shape rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_$tag.log | tail -1)"
  done
  timeout -k 10 400 compute-sanitizer --tool $tool --error-exitcode 9 python tools/profile_fused.py 1 4 48 64 > gpurun_out/sanitize_${tool}_fused_1_4_48_64.log 2>&1
  echo "$tool fused+stft 1 4 48 64 rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_fused_1_4_48_64.log | tail -1)"
done
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -k 10 900 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_final.log | cut -c1-300
timeout -k 10 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference_final.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_reference_final.log | cut -c1-200
timeout -k 10 300 python tools/shape_bench.py > gpurun_out/shape_bench_final.log 2>&1; echo "shape rc=$?"
