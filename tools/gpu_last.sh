#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --image-size 608 --batch 32 --no-cpu-baseline > gpurun_out/bench_608.log 2>&1; echo "bench 608 rc=$?"; tail -1 gpurun_out/bench_608.log | cut -c1-200
timeout 300 python tools/bench_detect.py > gpurun_out/bench_detect.log 2>&1; echo "detect rc=$?"; head -1 gpurun_out/bench_detect.log | cut -c1-330
timeout 300 python tools/run_layer.py L1 L2 L3 L4 L5 L6 L7 L8 L9 L10 L13 L14 L15 L19 L22 --iters 20 > gpurun_out/layers.log 2>&1; cat gpurun_out/layers.log
