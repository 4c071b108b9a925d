#!/bin/bash
# session 5: per-shape timing of the GROUPED pair launches (what the bench step is made of)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 300 python tools/shape_bench.py --what scan --pairs > gpurun_out/shape_pairs_s5k.log 2>&1; echo "rc=$?"; grep scan_ gpurun_out/shape_pairs_s5k.log | cut -c1-140 || tail -5 gpurun_out/shape_pairs_s5k.log
