#!/bin/bash
# session 6: cross scan / merge with the final tile rule (product library): parity, timing against the Triton kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_cross_gpu.py -x -q -m gpu --timeout 300 > gpurun_out/pytest_s6i.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_s6i.log
timeout -k 10 300 python tools/cross_vs_triton.py > gpurun_out/cross_vs_triton_final.log 2>&1; echo "== product rc=$?"
grep '"B"' gpurun_out/cross_vs_triton_final.log | python -c '
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d["C"], d["H"], d["W"], d["dtype"], "scan", d["cross_scan"]["ours_us"], d["cross_scan"]["speedup"], "merge", d["cross_merge"]["ours_us"], d["cross_merge"]["speedup"])
'
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
