#!/bin/bash
# Round 2, session m: epoch read / workspace recycling moved off the tile's critical path -- parity of everything, then timing.
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "gpu rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout -k 10 300 python tools/shape_bench.py > gpurun_out/shape_bench_m.log 2>&1; echo "shape rc=$?"; grep scan_ gpurun_out/shape_bench_m.log | cut -c1-130
timeout -k 10 600 python bench.py > gpurun_out/bench_m.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_m.log | cut -c1-900
