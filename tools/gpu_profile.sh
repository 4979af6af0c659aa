#!/bin/bash
# profiles for profiles/: launch list of one bench step + full capture of the dominant kernel + L1 source view
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 123 -c 41 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -f -o gpurun_out/ncu_full_L19 python tools/run_layer.py L19 --iters 1 > gpurun_out/ncu_full_L19.log 2>&1; echo "ncu L19 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -f -o gpurun_out/ncu_full_L1 python tools/run_layer.py L1 --iters 1 > gpurun_out/ncu_full_L1.log 2>&1; echo "ncu L1 rc=$?"
