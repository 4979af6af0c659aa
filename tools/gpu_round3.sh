#!/bin/bash
# round-1 final validation: GPU tests, smoke, both bench arms, launch list with DRAM bytes, full captures
# of the stream-K conv (L19), the halo conv (L2) and the fused decode+NMS kernel, microbenches
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=" gpurun_out/pytest_gpu.log | head -30
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-400
timeout 300 python tools/bench_detect.py > gpurun_out/bench_detect.log 2>&1; echo "detect rc=$?"; tail -3 gpurun_out/bench_detect.log
timeout 300 python tools/run_layer.py L1 L2 L3 L4 L5 L6 L7 L8 L9 L10 L13 L14 L15 L19 L22 --iters 20 > gpurun_out/layers.log 2>&1; cat gpurun_out/layers.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 100 -c 130 --csv --log-file gpurun_out/launches_infer.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv -s 1 -c 1 -f -o gpurun_out/ncu_full_L19 python tools/run_layer.py L19 --iters 1 --raw > gpurun_out/ncu_full_L19.log 2>&1; echo "ncu L19 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv -s 1 -c 1 -f -o gpurun_out/ncu_full_L2 python tools/run_layer.py L2 --iters 1 > gpurun_out/ncu_full_L2.log 2>&1; echo "ncu L2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:detect_fused -s 2 -c 1 -f -o gpurun_out/ncu_full_detect python tools/bench_detect.py > gpurun_out/ncu_full_detect.log 2>&1; echo "ncu detect rc=$?"
