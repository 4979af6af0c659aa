#!/bin/bash
# usage: gpu_round.sh [diag cases...]   -- full validation + bench + per-layer timing
mkdir -p gpurun_out
: > gpurun_out/diag_conv.log
for c in "$@"; do
  timeout 180 python tools/diag_conv.py $c >> gpurun_out/diag_conv.log 2>&1
  echo "case $c rc=$?" >> gpurun_out/diag_conv.log
done
grep -E '"case"|rc=' gpurun_out/diag_conv.log | cut -c1-300
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=|rel_l2" gpurun_out/pytest_gpu.log | head -40
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench.log
timeout 300 python tools/run_layer.py L1 L2 L3 L4 L5 L6 L7 L8 L9 L10 L13 L14 L15 L19 L22 > gpurun_out/layers.log 2>&1; cat gpurun_out/layers.log
