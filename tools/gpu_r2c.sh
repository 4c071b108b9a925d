#!/bin/bash
# Round 2, third GPU session: full parity suite, the rewritten bench line (both arms), ncu launch list of the bench step.
mkdir -p gpurun_out
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-200 | head -40
python -c "import json; d=json.load(open('gpurun_out/elementwise.json')); print(d['worst'])"
timeout -k 10 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.log 2>&1
echo "bench reference rc=$?"; tail -1 gpurun_out/bench_reference.log | cut -c1-300
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench.log
