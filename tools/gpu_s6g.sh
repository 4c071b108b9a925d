#!/bin/bash
# session 6: cross merge with the row-major pair's loads in front of the barrier, small maps with 1 / 2 / 4 planes per CTA
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_cross_gpu.py tests/test_ss2d_gpu.py -x -q -m gpu --timeout 300 > gpurun_out/pytest_s6g.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_s6g.log
timeout -k 10 600 python tools/cross_vs_triton.py > gpurun_out/cross_vs_triton_s6g.log 2>&1; echo "cross rc=$?"
grep '"B"' gpurun_out/cross_vs_triton_s6g.log | python -c '
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d["C"], d["H"], d["W"], d["dtype"], "scan", d["cross_scan"]["ours_us"], d["cross_scan"]["speedup"], "merge", d["cross_merge"]["ours_us"], d["cross_merge"]["speedup"])
'
