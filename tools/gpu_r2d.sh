#!/bin/bash
# Round 2, fourth GPU session: parity after the STFT rewrite / harness graph / prepared calls / drop-in test; bench line; Triton comparison.
mkdir -p gpurun_out
timeout -k 10 1800 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-220 | head -60
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench.log
timeout -k 10 600 python tools/cross_vs_triton.py > gpurun_out/cross_vs_triton.jsonl 2>gpurun_out/cross_vs_triton.err
echo "triton rc=$?"; cut -c1-500 gpurun_out/cross_vs_triton.jsonl; tail -3 gpurun_out/cross_vs_triton.err
