#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=" gpurun_out/pytest_gpu.log | head -30
timeout 600 python tools/bench_train.py --steps 10 > gpurun_out/bench_train_1.log 2>&1; echo "train1 rc=$?"; tail -1 gpurun_out/bench_train_1.log
timeout 600 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-400
