#!/bin/bash
# Round evidence run: smoke, parity tests, bench, ncu launch list of the bench command, full captures of the scan kernels,
# per-shape timing.
bash scripts_gpu_check.sh
bash tools/gpu_profile.sh
timeout -k 10 400 python tools/shape_bench.py --reps 20 > gpurun_out/shape_bench.log 2>&1
echo "shape bench rc=$?"; grep -c kernel gpurun_out/shape_bench.log
