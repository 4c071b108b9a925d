#!/bin/bash
# session 4, call a: the block-mode harness (whole SS2D bodies on the head / core / tail kernels) -- tests, then the bench line
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_ss2d_gpu.py tests/test_harness_gpu.py -x -q -k "block or harness or train_step or paired or graph or batched" --timeout 300 > gpurun_out/pytest_s4a.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_s4a.log
timeout -k 10 900 python bench.py > gpurun_out/bench_s4a.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_s4a.log | cut -c1-6000
