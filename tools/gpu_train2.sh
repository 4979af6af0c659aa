#!/bin/bash
# 2-GPU: DDP equivalence + training bench at 1 and 2 GPUs
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/ddp_check.log 2>&1; echo "ddp rc=$?"; tail -3 gpurun_out/ddp_check.log
timeout 600 python tools/bench_train.py --steps 10 > gpurun_out/bench_train_1.log 2>&1; echo "train1 rc=$?"; tail -2 gpurun_out/bench_train_1.log
timeout 600 python tools/bench_train.py --steps 10 --loss region > gpurun_out/bench_train_1r.log 2>&1; echo "train1r rc=$?"; tail -1 gpurun_out/bench_train_1r.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_train.py --steps 10 > gpurun_out/bench_train_2.log 2>&1; echo "train2 rc=$?"; tail -1 gpurun_out/bench_train_2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2.log 2>&1; echo "bench2 rc=$?"; tail -1 gpurun_out/bench_2.log
