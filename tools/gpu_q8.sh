#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x -k "bn_stats or engine or builders or training or trainer or affine" > gpurun_out/pytest_q8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_q8.log
grep -E "^E  |passed|failed|rc=|Error" gpurun_out/pytest_q8.log | head -20
for v in fused two fused two; do
  if [ $v = two ]; then export Y2_BN_STATS_TWO_KERNELS=1; else unset Y2_BN_STATS_TWO_KERNELS; fi
  timeout 600 python bench.py --no-cpu-baseline --steps 60 > gpurun_out/bench_$v.log 2>&1; echo "bench $v: $(tail -1 gpurun_out/bench_$v.log | cut -c70-110)"
done
unset Y2_BN_STATS_TWO_KERNELS
timeout 600 python tools/bench_train.py > gpurun_out/bench_train.log 2>&1; echo "train rc=$?"; tail -1 gpurun_out/bench_train.log | cut -c100-200
