#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_e2e_gpu.py -q -m gpu --timeout 600 > gpurun_out/pytest_e2e.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_e2e.log
grep -E "assert|Error|error|passed|failed|rc=" gpurun_out/pytest_e2e.log | head -40
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -5 gpurun_out/bench.log
