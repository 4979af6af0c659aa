#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 300 -x -k "conv1 or engine or builders_bf16" > gpurun_out/pytest_q10.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_q10.log
grep -E "^E  |passed|failed|rc=|Error" gpurun_out/pytest_q10.log | head -10
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline --steps 60 > gpurun_out/bench.log 2>&1; echo "bench: $(tail -1 gpurun_out/bench.log | cut -c70-110)"
