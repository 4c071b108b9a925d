#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_ss2d_gpu.py tests/test_model_dropin_gpu.py -x -q -k "conv_silu or block_core or drop_in or merge_norm_gate or core_out" --timeout 300 > gpurun_out/pytest_head.log 2>&1; echo "head+tail rc=$?"; tail -3 gpurun_out/pytest_head.log
rm -f gpurun_out/head_bench_xproj.jsonl gpurun_out/tail_bench.jsonl
for dt in float32 float16; do timeout -k 10 300 python tools/head_bench.py --dtype $dt --xproj 2>&1 | tee -a gpurun_out/head_bench_xproj.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('head', d['C'], d['dtype'], 'fwd', d['fused_fwd_us'], 'vs', d['chain_fwd_us'], 'fwd+bwd', d['fused_fwd_bwd_us'], 'vs', d['chain_fwd_bwd_us'])"; done
for dt in float32 float16; do timeout -k 10 300 python tools/tail_bench.py --dtype $dt 2>&1 | tee -a gpurun_out/tail_bench.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('tail', d['C'], d['dtype'], 'fwd', d['fused_fwd_us'], 'vs', d['chain_fwd_us'], 'fwd+bwd', d['fused_fwd_bwd_us'], 'vs', d['chain_fwd_bwd_us'])"; done
