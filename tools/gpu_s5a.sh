#!/bin/bash
# session 5, call a: state of the tree -- bench line, harness profile (kernels grouped by name), per-shape rows
mkdir -p gpurun_out
timeout -k 10 900 python bench.py > gpurun_out/bench_s5a.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_s5a.log | cut -c1-600
timeout -k 10 300 python tools/harness_profile.py --steps 2 > gpurun_out/harness_profile_s5a.txt 2>&1; echo "prof rc=$?"
timeout -k 10 300 python tools/shape_bench.py > gpurun_out/shape_bench_s5a.log 2>&1; echo "shape rc=$?"
