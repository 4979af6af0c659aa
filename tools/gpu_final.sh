#!/bin/bash
# final check of the committed state: GPU tests, smoke, both bench arms, training bench, launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=" gpurun_out/pytest_gpu.log | head -30
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-260
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-200
timeout 600 python tools/bench_train.py > gpurun_out/bench_train.log 2>&1; echo "train rc=$?"; tail -1 gpurun_out/bench_train.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 90 -c 120 --csv --log-file gpurun_out/launches_infer.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
