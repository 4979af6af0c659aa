#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_e2e_gpu.py -q -k "detect or engine" --timeout 300 > gpurun_out/pytest_detect.log 2>&1; echo "detect tests rc=$?"; grep -E "^E  |passed|failed" gpurun_out/pytest_detect.log | head
timeout 300 python tools/bench_detect.py 2>&1 | tail -2
timeout 300 python tools/bench_detect.py --grid 19 2>&1 | head -1
