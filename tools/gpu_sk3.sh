#!/bin/bash
mkdir -p gpurun_out
echo "--- default"; timeout 120 python tools/run_layer.py L6 L7 L10 L15 --iters 20 2>&1 | tail -4
echo "--- min_ksteps 4 (2cta)"; Y2_CONV_STREAMK_MIN_KSTEPS=4 timeout 120 python tools/run_layer.py L6 L10 L15 --iters 20 2>&1 | tail -3
echo "--- min_ksteps 4 (1cta)"; Y2_CONV_STREAMK_1CTA=1 Y2_CONV_STREAMK_MIN_KSTEPS=4 timeout 120 python tools/run_layer.py L6 L10 L15 --iters 20 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv -s 1 -c 1 -f -o gpurun_out/ncu_full_L19_2cta python tools/run_layer.py L19 --iters 1 --raw > gpurun_out/ncu_full_L19_2cta.log 2>&1; echo "ncu rc=$?"
Y2_CONV_STREAMK_MIN_KSTEPS=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv -s 1 -c 1 -f -o gpurun_out/ncu_full_L10_2cta python tools/run_layer.py L10 --iters 1 > gpurun_out/ncu_full_L10_2cta.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
