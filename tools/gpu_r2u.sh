#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_ss2d_gpu.py tests/test_model_dropin_gpu.py -x -q -k "conv_silu or block_core or drop_in" --timeout 300 > gpurun_out/pytest_head.log 2>&1; echo "head rc=$?"; tail -30 gpurun_out/pytest_head.log
