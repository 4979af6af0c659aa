#!/bin/bash
mkdir -p gpurun_out
export Y2_CONV_TMA_STORE=1
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x -k "halo_pair or engine_full_size or builders_bf16 or passthrough" > gpurun_out/pytest_t3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_t3.log
grep -E "^E  |passed|failed|rc=|y2 conv|Error" gpurun_out/pytest_t3.log | head -20
echo "--- tma store"; timeout 120 python tools/run_layer.py L3 --iters 20 2>&1 | tail -1
echo "--- direct stores"; Y2_CONV_NO_TMA_STORE=1 timeout 120 python tools/run_layer.py L3 --iters 20 2>&1 | tail -1
echo "--- tma store"; timeout 120 python tools/run_layer.py L3 --iters 20 2>&1 | tail -1
echo "--- direct stores"; Y2_CONV_NO_TMA_STORE=1 timeout 120 python tools/run_layer.py L3 --iters 20 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench tma rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-200
Y2_CONV_NO_TMA_STORE=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_direct.log 2>&1; echo "bench direct rc=$?"; tail -1 gpurun_out/bench_direct.log | cut -c1-200
