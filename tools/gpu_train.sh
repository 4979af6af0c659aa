#!/bin/bash
# training-path bring-up: parity tests of the backward kernels and the full training step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_training_gpu.py -q -m gpu --timeout 300 -s "$@" > gpurun_out/pytest_train.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_train.log
grep -E "^E  |passed|failed|rc=|rel_l2|PASS|FAIL|Error|error" gpurun_out/pytest_train.log | head -80
