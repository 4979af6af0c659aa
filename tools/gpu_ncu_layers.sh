#!/bin/bash
mkdir -p gpurun_out
for L in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 1 -c 1 -f -o gpurun_out/ncu_full_$L python tools/run_layer.py $L --iters 1 > gpurun_out/ncu_full_$L.log 2>&1
  echo "ncu $L rc=$?"
done
