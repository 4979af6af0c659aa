#!/bin/bash
# profiles for profiles/: launch list (durations + DRAM bytes) of bench steps, full captures of the top kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 100 -c 130 --csv --log-file gpurun_out/launches_infer.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -f -o gpurun_out/ncu_full_L19 python tools/run_layer.py L19 --iters 1 > gpurun_out/ncu_full_L19.log 2>&1; echo "ncu L19 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1_u8 -s 2 -c 1 -f -o gpurun_out/ncu_full_conv1 python tools/bench_conv1.py > gpurun_out/ncu_full_conv1.log 2>&1; echo "ncu conv1 rc=$?"
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300
