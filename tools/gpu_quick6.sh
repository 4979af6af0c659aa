#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=" gpurun_out/pytest_gpu.log | head -30
for impl in split fused; do
  Y2_DETECT_IMPL=$impl timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$impl.log 2>&1; echo "bench $impl rc=$?"; tail -1 gpurun_out/bench_$impl.log | cut -c1-260
done
