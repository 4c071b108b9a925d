#!/bin/bash
# Round 2, seventh GPU session: channels-per-tile experiment of the multi-chunk backward (tuning build), harness test, launch list + full captures for profiles/.
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_harness_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_harness.log 2>&1
echo "harness tests rc=$?"; tail -2 gpurun_out/pytest_harness.log
export VMASR_B200_LIBRARY=$PWD/vm_asr_b200/lib_tuning/libvmasr_b200.so
for cpt in 4 3 2; do
  VMASR_SCAN_CPT=$cpt timeout -k 10 200 python tools/shape_bench.py --reps 20 --what scan > gpurun_out/shape_bench_cpt$cpt.log 2>&1
  echo "bwd channels per tile $cpt:"; grep scan_bwd gpurun_out/shape_bench_cpt$cpt.log | cut -c1-120
done
for cpt in 3 2; do
  VMASR_SCAN_CPT_FWD=$cpt timeout -k 10 200 python tools/shape_bench.py --reps 20 --what scan > gpurun_out/shape_bench_fcpt$cpt.log 2>&1
  echo "fwd channels per tile $cpt:"; grep scan_fwd gpurun_out/shape_bench_fcpt$cpt.log | cut -c1-120
done
unset VMASR_B200_LIBRARY
