#!/bin/bash
# session 5: after the bar_par change -- sanitizer again (with a multi-tile shape), the grouped-launch tests, the bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_scan_gpu.py -m gpu -q -x -k "grouped or config_shapes or graph" --timeout 120 --timeout-method=thread > gpurun_out/pytest_s5i.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_s5i.log
bash tools/gpu_sanitize.sh
timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline > gpurun_out/bench_s5i.log 2>&1; echo "bench (product) rc=$?"; tail -1 gpurun_out/bench_s5i.log | cut -c1-200
