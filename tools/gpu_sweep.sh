#!/bin/bash
# Tuning sweep of the scan fast path: ring depth and channels per tile, per-shape timing each.  Run under gpurun.
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout -k 10 300 python tools/shape_bench.py --reps 20 --what scan > gpurun_out/sweep_$tag.log 2>&1
  echo "== $tag ($*) rc=$?"
  grep scan_ gpurun_out/sweep_$tag.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('  %-9s D=%-5d L=%-7d %8.4f ms %7.1f GB/s %.3f' % (r['kernel'], r['D'], r['L'], r['ms'], r['GBps'], r['frac']))"
}
for cfg in "$@"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  run "$tag" $cfg
done
