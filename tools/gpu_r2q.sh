#!/bin/bash
# Round 2, session q: merge + LayerNorm + gate kernels -- parity, then the model-level drop-in with the fused tail.
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_ss2d_gpu.py -x -q -k "merge_norm or core_out" --timeout 300 > gpurun_out/pytest_tail.log 2>&1; echo "tail rc=$?"; tail -25 gpurun_out/pytest_tail.log
timeout -k 10 600 python -m pytest tests/test_model_dropin_gpu.py tests/test_harness_gpu.py -x -q --timeout 300 > gpurun_out/pytest_dropin.log 2>&1; echo "dropin rc=$?"; tail -15 gpurun_out/pytest_dropin.log
