#!/bin/bash
# Round 2, session o: backward tiles in rounds -- parity, then channels-per-tile sweep (tuning build) and the default plan.
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_scan_gpu.py tests/test_ss2d_gpu.py tests/test_vs_reference_cuda_gpu.py -x -q --timeout 600 > gpurun_out/pytest_rounds.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_rounds.log
timeout -k 10 300 python tools/shape_bench.py > gpurun_out/shape_bench_o.log 2>&1; echo "shape rc=$?"; grep scan_bwd gpurun_out/shape_bench_o.log | cut -c1-130
for cpt in 4 8 12 16; do
  echo "== VMASR_SCAN_CPT=$cpt"
  VMASR_B200_LIBRARY=$PWD/vm_asr_b200/lib_tuning/libvmasr_b200.so VMASR_SCAN_CPT=$cpt timeout -k 10 300 python tools/shape_bench.py 2>&1 | grep scan_bwd | grep -v "L\": 1024\|L\": 256," | cut -c1-130 | tee -a gpurun_out/shape_bench_o_cpt$cpt.log
done
timeout -k 10 600 python bench.py --steps 30 > gpurun_out/bench_o.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_o.log | cut -c1-300
