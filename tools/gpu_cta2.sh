#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=|y2 conv|Error" gpurun_out/pytest_gpu.log | head -30
echo "--- cta2"; timeout 120 python tools/run_layer.py L2 L3 L4 L5 --iters 20 2>&1 | tail -4
echo "--- no cta2"; Y2_CONV_NO_CTA2=1 timeout 120 python tools/run_layer.py L2 L3 L4 L5 --iters 20 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-200
