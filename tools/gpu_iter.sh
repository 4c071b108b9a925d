#!/bin/bash
# One kernel-tuning iteration on the GPU: scan parity tests, dynamic instruction counts, per-shape timing.
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_scan_gpu.py -m gpu -q -x --timeout 120 --timeout-method=thread > gpurun_out/pytest_scan.log 2>&1
echo "pytest scan rc=$?"; tail -3 gpurun_out/pytest_scan.log
bash tools/gpu_inst.sh
