#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=|y2 conv|Error" gpurun_out/pytest_gpu.log | head -30
LAYERS="L7 L8 L10 L13 L15 L22"
echo "--- pairs"; timeout 120 python tools/run_layer.py $LAYERS --iters 20 2>&1 | tail -6
echo "--- no generic pairs"; Y2_CONV_NO_CTA2_GENERIC=1 timeout 120 python tools/run_layer.py $LAYERS --iters 20 2>&1 | tail -6
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-200
timeout 600 python tools/bench_train.py > gpurun_out/bench_train.log 2>&1; echo "train rc=$?"; tail -1 gpurun_out/bench_train.log | cut -c1-400
