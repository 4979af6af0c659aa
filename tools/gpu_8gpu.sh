#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_8.log 2>&1; echo "bench8 rc=$?"; tail -1 gpurun_out/bench_8.log | cut -c1-300
