#!/bin/bash
# Round 2, sixth GPU session: parity and bench after the TMA tensor-map / swizzle conversion, STFT fast epilogues, harness AMP.
mkdir -p gpurun_out
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -k 10 1800 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-220 | head -60
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench.log
timeout -k 10 300 python tools/shape_bench.py --reps 20 --what scan > gpurun_out/shape_bench.log 2>&1
echo "shape bench rc=$?"; cat gpurun_out/shape_bench.log | cut -c1-200
timeout -k 10 300 python tools/harness_profile.py > gpurun_out/harness_profile.txt 2>&1
echo "harness profile rc=$?"; sed -n 4,30p gpurun_out/harness_profile.txt | cut -c1-230
