#!/bin/bash
# Round 2, session i: whole GPU suite on poisoned allocator memory (tests/conftest.py), harness test repeated.
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/pytest_poison.log 2>&1
echo "pytest poisoned rc=$?"
tail -40 gpurun_out/pytest_poison.log
for i in 1 2 3; do
  timeout -k 10 300 python -m pytest tests/test_harness_gpu.py -q --timeout 300 -x > gpurun_out/pytest_harness_$i.log 2>&1
  echo "harness run $i rc=$?"; tail -3 gpurun_out/pytest_harness_$i.log
done
VMASR_NO_POISON=1 timeout -k 10 300 python -m pytest tests/test_harness_gpu.py -q --timeout 300 -x > gpurun_out/pytest_harness_nopoison.log 2>&1
echo "harness unpoisoned rc=$?"; tail -3 gpurun_out/pytest_harness_nopoison.log
