#!/bin/bash
# session 6: cross scan / merge tile size -- 64 x 64 tiles always (VMASR_CROSS_T32=0) vs 32 x 32 tiles always (=1000000000), tuning build;
# then the product rule; parity of the cross ops
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
show() { grep '"B"' $1 | python -c '
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d["C"], d["H"], d["W"], d["dtype"], "scan", d["cross_scan"]["ours_us"], d["cross_scan"]["speedup"], "merge", d["cross_merge"]["ours_us"], d["cross_merge"]["speedup"])
'; }
for t in 0 1000000000; do
  VMASR_CROSS_T32=$t VMASR_B200_LIBRARY=vm_asr_b200/lib_tuning/libvmasr_b200.so timeout -k 10 300 python tools/cross_vs_triton.py > gpurun_out/cross_vs_triton_t32_$t.log 2>&1; echo "== VMASR_CROSS_T32=$t rc=$?"; show gpurun_out/cross_vs_triton_t32_$t.log
done
timeout -k 10 300 python tools/cross_vs_triton.py > gpurun_out/cross_vs_triton_final.log 2>&1; echo "== product rc=$?"; show gpurun_out/cross_vs_triton_final.log
timeout -k 10 600 python -m pytest tests/test_cross_gpu.py -x -q -m gpu --timeout 300 > gpurun_out/pytest_s6h.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_s6h.log
