#!/bin/bash
# session 5, records with the final library (programmatic dependent launch on): smoke, bench line, reference arm, per-shape rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -k 10 900 python bench.py > gpurun_out/bench_final.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_final.log | cut -c1-300
timeout -k 10 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference_final.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_reference_final.log | cut -c1-200
timeout -k 10 300 python tools/shape_bench.py > gpurun_out/shape_bench_final.log 2>&1; echo "shape rc=$?"
timeout -k 10 300 python tools/shape_bench.py --what scan --pairs > gpurun_out/shape_pairs_final.log 2>&1; echo "pairs rc=$?"; grep scan_ gpurun_out/shape_pairs_final.log | cut -c1-130
timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline --workload vm_asr_48k_16k_nfft2048 > gpurun_out/bench_nfft2048.log 2>&1; echo "nfft2048 rc=$?"; tail -1 gpurun_out/bench_nfft2048.log | cut -c1-200
timeout -k 10 600 python bench.py --no-e2e --no-core --no-stft --no-cpu-baseline --workload vm_asr_48k_16k_MPD_VSSM32 > gpurun_out/bench_vssm32.log 2>&1; echo "vssm32 rc=$?"; tail -1 gpurun_out/bench_vssm32.log | cut -c1-200
