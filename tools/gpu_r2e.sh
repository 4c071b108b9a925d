#!/bin/bash
# Round 2, fifth GPU session: where the harness step's time goes; full ncu captures (source level) of the STFT kernels and of the scan kernels.
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_cross_gpu.py tests/test_stft_gpu.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_fix.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_fix.log
timeout -k 10 300 python tools/harness_profile.py > gpurun_out/harness_profile.txt 2>&1
echo "harness profile rc=$?"; head -60 gpurun_out/harness_profile.txt | cut -c1-230
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"stft|synth|finalize|istft" -s 5 -c 5 -f -o gpurun_out/prof_stft python tools/profile_stft.py 4 > gpurun_out/prof_stft.log 2>&1
echo "ncu stft rc=$?"
bash tools/gpu_ncu_one.sh "4 64 65536" r2_4_64_65536
bash tools/gpu_ncu_one.sh "4 256 4096" r2_4_256_4096
ls -la gpurun_out/*.ncu-rep
