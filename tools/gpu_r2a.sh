#!/bin/bash
# Round 2, first GPU session: parity of the reworked scan kernels (reverse / accumulate / grouped), the fused SS2D core and the
# rewritten cross kernels; bench with the three pairing modes; fused-vs-chain per shape.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
for mode in grouped streams none; do
  timeout -k 10 300 python bench.py --steps 20 --warmup 5 --pairing $mode --no-e2e --no-cpu-baseline > gpurun_out/bench_$mode.log 2>&1
  echo "bench $mode rc=$?"; tail -1 gpurun_out/bench_$mode.log | cut -c1-400
done
timeout -k 10 300 python tools/ss2d_bench.py --reps 10 > gpurun_out/ss2d_bench.log 2>&1
echo "ss2d bench rc=$?"; cat gpurun_out/ss2d_bench.log | cut -c1-600
timeout -k 10 300 python tools/ss2d_bench.py --reps 10 --pair > gpurun_out/ss2d_bench_pair.log 2>&1
echo "ss2d bench pair rc=$?"; cat gpurun_out/ss2d_bench_pair.log | cut -c1-600
timeout -k 10 300 python tools/shape_bench.py --reps 20 --what scan,cross > gpurun_out/shape_bench.log 2>&1
echo "shape bench rc=$?"; cat gpurun_out/shape_bench.log | cut -c1-300
