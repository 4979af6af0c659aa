#!/bin/bash
# A/B: library variants on the same box, alternating.  usage: gpu_ab.sh "libA libB ..." L1 L2 ...
mkdir -p gpurun_out
LIBS="$1"; shift
for i in 1 2; do
  for lib in $LIBS; do
    echo "--- $lib $i"; Y2_LIB_PATH=$PWD/tensorflow_yolo2_b200/lib/$lib python tools/run_layer.py "$@" --iters 20 2>&1 | grep -E "TFLOP|rror"
  done
done
