#!/bin/bash
# Round 2, session k: harness stability check after the pre-normalisation fix; phase timelines of the multi-chunk scan kernels.
mkdir -p gpurun_out
LR=2e-4 timeout -k 10 300 python tools/debug_harness.py > gpurun_out/debug_harness.log 2>&1; echo "debug rc=$?"
grep -c "bad_grads 0" gpurun_out/debug_harness.log; tail -12 gpurun_out/debug_harness.log
for i in 1 2 3; do timeout -k 10 300 python -m pytest tests/test_harness_gpu.py -q --timeout 300 2>&1 | tail -2; done
export VMASR_B200_LIBRARY=$PWD/vm_asr_b200/lib_tuning/libvmasr_b200.so
for shape in "4 64 65536" "4 128 16384" "4 256 4096" "4 8 262144"; do
  timeout -k 10 120 python tools/timeline.py $shape 2>&1 | tee -a gpurun_out/timeline.log
done
