#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "tests rc=$?"; grep -E "^E  |passed|failed|rel_l2" gpurun_out/pytest_gpu.log | head -20
timeout 600 python bench.py --no-cpu-baseline --steps 50 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-400
