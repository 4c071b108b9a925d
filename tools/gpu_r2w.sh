#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/head_bench_xproj.jsonl
for dt in float32 float16; do timeout -k 10 300 python tools/head_bench.py --dtype $dt --xproj 2>&1 | tee -a gpurun_out/head_bench_xproj.jsonl | cut -c1-330; done
