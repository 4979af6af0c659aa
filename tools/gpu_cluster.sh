#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/diag_conv.log
for c in i1x1_64_128 t1x1_64_128 i3x3_64_128_odd i3x3_128_64 i3x3_32_64 t3x3_32_64_pool t3x3_64_128_pool t3x3_256_512_pool26 i3x3_512_1024 i1x1_1024_512 i1x1_1024_125_f32 i1x1_1024_30_f32 i3x3_1024_1024_f32 p3x3_256_512 p3x3_128_64 p3x3_3_32_pool big_l3; do
  timeout 120 python tools/diag_conv.py $c >> gpurun_out/diag_conv.log 2>&1
  echo "case $c rc=$?" >> gpurun_out/diag_conv.log
done
grep -E '"case"|rc=' gpurun_out/diag_conv.log | cut -c1-200
for i in 1 2; do
echo "--- old"; Y2_LIB_PATH=$PWD/tensorflow_yolo2_b200/lib/libyolo2_old.so python tools/run_layer.py L6 L9 L14 L19 L15 L10 L7 --iters 20 2>&1 | grep -E "TFLOP|rror"
echo "--- nocluster"; Y2_CONV_NO_CLUSTER=1 python tools/run_layer.py L6 L9 L14 L19 L15 L10 L7 --iters 20 2>&1 | grep -E "TFLOP|rror"
echo "--- cluster"; python tools/run_layer.py L6 L9 L14 L19 L15 L10 L7 --iters 20 2>&1 | grep -E "TFLOP|rror"
done
