#!/bin/bash
# 2-GPU closing run: DDP equivalence, inference bench and training bench at 2 GPUs
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/ddp_check.log 2>&1; echo "ddp rc=$?"; tail -3 gpurun_out/ddp_check.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_2.log 2>&1; echo "bench2 rc=$?"; tail -1 gpurun_out/bench_2.log | cut -c1-700
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_train.py --steps 10 > gpurun_out/bench_train_2.log 2>&1; echo "train2 rc=$?"; tail -1 gpurun_out/bench_train_2.log | cut -c1-600
