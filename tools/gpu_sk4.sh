#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 200 -x -k "streamk" > gpurun_out/pytest_sk.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_sk.log
grep -E "^E  |passed|failed|rc=|y2 conv" gpurun_out/pytest_sk.log | head -30
echo "--- 512-row pair tiles"; timeout 120 python tools/run_layer.py L6 L9 L14 L19 --iters 20 2>&1 | tail -4
echo "--- raw"; timeout 120 python tools/run_layer.py L14 L19 --raw --iters 20 2>&1 | tail -2
echo "--- 256-row pair tiles"; Y2_CONV_STREAMK_256=1 timeout 120 python tools/run_layer.py L6 L9 L14 L19 --iters 20 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_sk512.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_sk512.log | cut -c1-200
