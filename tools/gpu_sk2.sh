#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 120 -x -k "streamk" > gpurun_out/pytest_sk.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_sk.log
grep -E "^E  |passed|failed|rc=|y2 conv" gpurun_out/pytest_sk.log | head -30
echo "--- 2cta"; timeout 120 python tools/run_layer.py L9 L14 L19 --iters 20 2>&1 | tail -4
echo "--- 2cta raw"; timeout 120 python tools/run_layer.py L14 L19 --raw --iters 20 2>&1 | tail -3
echo "--- 1cta"; Y2_CONV_STREAMK_1CTA=1 timeout 120 python tools/run_layer.py L9 L14 L19 --iters 20 2>&1 | tail -4
