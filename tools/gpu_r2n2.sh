#!/bin/bash
# 2-GPU check of the bench line (NCCL: flat bucketed gradient all-reduce + payload, overlapped) and of the new GPU tests.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_cross_gpu.py tests/test_harness_gpu.py -m gpu -q --timeout 300 > gpurun_out/pytest_new.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_new.log
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.log 2>&1
echo "bench n2 rc=$?"; tail -1 gpurun_out/bench_n2.log
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/bench_ref_n2.log 2>&1
echo "bench ref n2 rc=$?"; tail -1 gpurun_out/bench_ref_n2.log | cut -c1-300
