#!/bin/bash
# Round 2, session l: delta generated inside the scan kernels (projected form) -- parity, then the whole GPU suite, then timing.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_ss2d_gpu.py -x -q -k "projected" --timeout 300 > gpurun_out/pytest_proj.log 2>&1; echo "proj rc=$?"; tail -15 gpurun_out/pytest_proj.log
timeout -k 10 1500 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "gpu rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout -k 10 400 python tools/ss2d_bench.py --projected --only 2,16,32 > gpurun_out/ss2d_bench_proj.jsonl 2> gpurun_out/ss2d_bench_proj.err; echo "bench rc=$?"; cat gpurun_out/ss2d_bench_proj.jsonl; tail -3 gpurun_out/ss2d_bench_proj.err
