#!/bin/bash
# full ncu captures of single layers (source-level stall sampling -> which role waits on which barrier)
mkdir -p gpurun_out
for L in ${LAYERS:-L2 L3 L4 L6 L7}; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv -s 1 -c 1 -f -o gpurun_out/ncu_full_$L python tools/run_layer.py $L --iters 1 $EXTRA > gpurun_out/ncu_full_$L.log 2>&1; echo "ncu $L rc=$?"
done
