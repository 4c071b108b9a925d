#!/bin/bash
# 2 GPUs of one box: the bench line as the driver launches it (torchrun, one rank per GPU over NCCL), both arms.
mkdir -p gpurun_out
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.log 2>&1; echo "n2 rc=$?"; tail -1 gpurun_out/bench_n2.log | cut -c1-300
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/bench_ref_n2.log 2>&1; echo "ref n2 rc=$?"; tail -1 gpurun_out/bench_ref_n2.log | cut -c1-200
