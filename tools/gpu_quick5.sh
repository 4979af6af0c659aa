#!/bin/bash
# tile pairing A/B: GPU tests, per-layer timings with / without pairing, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=" gpurun_out/pytest_gpu.log | head -30
LAYERS="L2 L3 L4 L5 L6 L7 L8 L9 L10 L13 L14 L15 L22"
echo "--- paired"; timeout 300 python tools/run_layer.py $LAYERS --iters 20 2>&1 | tee gpurun_out/layers_pair.log
echo "--- no pair"; Y2_CONV_NO_PAIR=1 timeout 300 python tools/run_layer.py $LAYERS --iters 20 2>&1 | tee gpurun_out/layers_nopair.log
echo "--- no pair256"; Y2_CONV_NO_PAIR256=1 timeout 300 python tools/run_layer.py L6 L8 L10 L13 L15 --iters 20 2>&1 | tee gpurun_out/layers_nopair256.log
echo "--- pair, block_n 128 (no streamk)"; Y2_CONV_NO_STREAMK=1 Y2_CONV_BLOCK_N=128 timeout 300 python tools/run_layer.py L6 L8 L9 L10 L13 L14 L15 --iters 20 2>&1 | tee gpurun_out/layers_bn128.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
