#!/bin/bash
# re-entry validation: GPU tests, smoke, bench (both arms), decode+NMS microbench, training bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=" gpurun_out/pytest_gpu.log | head -30
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 gpurun_out/bench_ref.log
timeout 300 python tools/bench_detect.py > gpurun_out/bench_detect.log 2>&1; echo "detect rc=$?"; tail -2 gpurun_out/bench_detect.log
timeout 600 python tools/bench_train.py --steps 10 > gpurun_out/bench_train_1.log 2>&1; echo "train1 rc=$?"; tail -1 gpurun_out/bench_train_1.log
