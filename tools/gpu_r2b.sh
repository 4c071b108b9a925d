#!/bin/bash
# Round 2, second GPU session: parity after the workspace-layout fix; fused core two-pass (default) vs one-pass (tuning build);
# ncu launch list of the fused and chained core on two shapes.
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-200 | head -60
cat gpurun_out/elementwise.json | head -12
timeout -k 10 300 python tools/ss2d_bench.py --reps 10 > gpurun_out/ss2d_bench.log 2>&1
echo "ss2d bench (two passes) rc=$?"; cut -c1-700 gpurun_out/ss2d_bench.log
timeout -k 10 300 python tools/ss2d_bench.py --reps 10 --pair > gpurun_out/ss2d_bench_pair.log 2>&1
echo "ss2d bench pair (two passes) rc=$?"; cut -c1-700 gpurun_out/ss2d_bench_pair.log
export VMASR_B200_LIBRARY=$PWD/vm_asr_b200/lib_tuning/libvmasr_b200.so
VMASR_SS2D_ONE_PASS=1 timeout -k 10 300 python tools/ss2d_bench.py --reps 10 --pair > gpurun_out/ss2d_bench_pair_onepass.log 2>&1
echo "ss2d bench pair (one pass, red.add) rc=$?"; cut -c1-700 gpurun_out/ss2d_bench_pair_onepass.log
unset VMASR_B200_LIBRARY
timeout -k 10 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
   --log-file gpurun_out/ss2d_launches.csv python tools/ss2d_bench.py --reps 2 --only 32,2 > gpurun_out/ss2d_ncu.log 2>&1
echo "ncu rc=$?"; python tools/launch_list.py gpurun_out/ss2d_launches.csv gpurun_out/ss2d_launches_ours.csv gpurun_out/ss2d_traffic.json; tail -25 gpurun_out/ss2d_launches_ours.csv
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300
