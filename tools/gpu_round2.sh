#!/bin/bash
# full validation + bench + DRAM-traffic launch lists (inference step, training step)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=" gpurun_out/pytest_gpu.log | head -30
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 gpurun_out/bench_ref.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 123 -c 41 --csv --log-file gpurun_out/launches_infer.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "infer launch list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 864 -c 288 --csv --log-file gpurun_out/launches_train.csv python tools/bench_train.py --steps 1 --warmup 3 > gpurun_out/train_under_ncu.log 2>&1
echo "train launch list rc=$?"
