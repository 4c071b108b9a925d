#!/bin/bash
# Round 2, session p: mbarrier waits with a suspend-time hint (NANOSLEEP.SYNCS instead of polling).
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_scan_gpu.py tests/test_ss2d_gpu.py -x -q --timeout 600 > gpurun_out/pytest_p.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_p.log
timeout -k 10 300 python tools/shape_bench.py > gpurun_out/shape_bench_p.log 2>&1; echo "shape rc=$?"; grep scan_ gpurun_out/shape_bench_p.log | cut -c1-130
timeout -k 10 600 python bench.py --steps 30 > gpurun_out/bench_p.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_p.log | cut -c1-300
