#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 -x -k "detect" > gpurun_out/pytest_det.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_det.log
grep -E "^E  |passed|failed|rc=|y2 detect|Error" gpurun_out/pytest_det.log | head -30
timeout 300 python tools/bench_detect.py 2>&1 | tee gpurun_out/bench_detect.log | tail -3
