#!/bin/bash
mkdir -p gpurun_out
run() { echo "--- $1"; env $1 python tools/run_layer.py L1 L2 L3 L4 L5 --iters 20 2>&1 | grep -E "TFLOP|rror"; }
for i in 1 2; do
run "X=1"
run "Y2_CONV_NO_PATCH=1"
run "Y2_CONV_NO_BSTAT=1"
run "Y2_CONV_NO_PATCH=1 Y2_CONV_NO_BSTAT=1"
run "Y2_CONV_NO_CLUSTER=1"
done
