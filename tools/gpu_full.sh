#!/bin/bash
# Full GPU session: smoke, all parity tests, bench, per-shape timing.
bash scripts_gpu_check.sh
timeout -k 10 600 python tools/shape_bench.py --reps 20 > gpurun_out/shape_bench.log 2>&1
echo "shape bench rc=$?"; grep scan_ gpurun_out/shape_bench.log
