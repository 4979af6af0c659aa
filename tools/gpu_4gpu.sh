#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_4.log 2>&1; echo "bench4 rc=$?"; tail -1 gpurun_out/bench_4.log | cut -c1-300
