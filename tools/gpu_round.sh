#!/bin/bash
# Round evidence run: smoke, parity tests, bench, ncu launch list of the bench command, full captures of the scan kernels.
bash scripts_gpu_check.sh
bash tools/gpu_profile.sh
