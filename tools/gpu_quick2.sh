#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=|^FAILED" gpurun_out/pytest_gpu.log | head -30
timeout 300 python tools/bench_detect.py > gpurun_out/bench_detect.log 2>&1; echo "detect rc=$?"; cat gpurun_out/bench_detect.log | cut -c1-600
timeout 600 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300
