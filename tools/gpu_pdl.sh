#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=|Error" gpurun_out/pytest_gpu.log | head -30
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pdl.log 2>&1; echo "bench pdl rc=$?"; tail -1 gpurun_out/bench_pdl.log | cut -c1-200
Y2_NO_PDL=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_nopdl.log 2>&1; echo "bench nopdl rc=$?"; tail -1 gpurun_out/bench_nopdl.log | cut -c1-200
timeout 600 python tools/bench_train.py > gpurun_out/bench_train.log 2>&1; echo "train rc=$?"; tail -1 gpurun_out/bench_train.log | cut -c1-300
