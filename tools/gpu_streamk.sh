#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "streamk" --timeout 300 > gpurun_out/pytest_sk.log 2>&1; echo "streamk tests rc=$?"; grep -E "^E  |passed|failed|timeout|never" gpurun_out/pytest_sk.log | head
echo "--- generic raw"; Y2_CONV_NO_STREAMK=1 timeout 120 python tools/run_layer.py L14 L19 --iters 10 --raw
echo "--- streamk raw"; timeout 120 python tools/run_layer.py L14 L19 --iters 10 --raw
echo "--- generic fused"; Y2_CONV_NO_STREAMK=1 timeout 120 python tools/run_layer.py L9 L11 L14 L16 --iters 10
echo "--- streamk fused"; timeout 120 python tools/run_layer.py L9 L11 L14 L16 --iters 10
