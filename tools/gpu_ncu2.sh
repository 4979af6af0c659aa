#!/bin/bash
# ncu --set full of one layer with several library variants. usage: gpu_ncu2.sh "libA libB" L19
mkdir -p gpurun_out
LIBS="$1"; shift
for lib in $LIBS; do
  for L in "$@"; do
    tag=$(basename $lib .so)_$L
    Y2_LIB_PATH=$PWD/tensorflow_yolo2_b200/lib/$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -f -o gpurun_out/ncu_$tag python tools/run_layer.py $L --iters 1 > gpurun_out/ncu_$tag.log 2>&1
    echo "ncu $tag rc=$?"
  done
done
