#!/bin/bash
# quick loop: GPU tests, per-layer timings, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=" gpurun_out/pytest_gpu.log | head -30
timeout 300 python tools/run_layer.py L1 L2 L3 L4 L5 L6 L7 L8 L9 L10 L13 L14 L15 L19 L22 --iters 20 > gpurun_out/layers.log 2>&1; cat gpurun_out/layers.log
timeout 300 python tools/run_layer.py L14 L19 --raw --iters 20 2>&1 | tee gpurun_out/layers_raw.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
