#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py tests/test_kernels_gpu.py -q -m gpu --timeout 300 -x -k "bn_bwd or training or affine or trainer or step or loss" > gpurun_out/pytest_train.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_train.log
grep -E "^E  |passed|failed|rc=|Error" gpurun_out/pytest_train.log | head -20
timeout 600 python tools/bench_train.py > gpurun_out/bench_train.log 2>&1; echo "train rc=$?"; tail -1 gpurun_out/bench_train.log | cut -c1-420
