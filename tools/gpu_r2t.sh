#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_ss2d_gpu.py tests/test_model_dropin_gpu.py -x -q -k "merge_norm or core_out or drop_in" --timeout 300 > gpurun_out/pytest_tail.log 2>&1; echo "tail rc=$?"; tail -4 gpurun_out/pytest_tail.log
rm -f gpurun_out/tail_bench.jsonl
for dt in float32 float16; do timeout -k 10 300 python tools/tail_bench.py --dtype $dt 2>&1 | tee -a gpurun_out/tail_bench.jsonl | cut -c1-420; done
