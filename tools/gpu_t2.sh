#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x -k "classifier or generic_pair or affine" -s > gpurun_out/pytest_t2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_t2.log
grep -E "^E  |passed|failed|rc=|rel_l2|Error" gpurun_out/pytest_t2.log | head -30
