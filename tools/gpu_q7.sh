#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x -k "bn_stats or engine or builders_bf16 or training_step or trainer" > gpurun_out/pytest_q7.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_q7.log
grep -E "^E  |passed|failed|rc=|Error" gpurun_out/pytest_q7.log | head -20
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-200
timeout 600 python tools/bench_train.py > gpurun_out/bench_train.log 2>&1; echo "train rc=$?"; tail -1 gpurun_out/bench_train.log | cut -c1-300
