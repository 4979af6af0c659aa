#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=|Error" gpurun_out/pytest_gpu.log | head -20
for v in pdl off pdl off; do
  if [ $v = off ]; then export Y2_NO_PDL_SMALL=1; else unset Y2_NO_PDL_SMALL; fi
  timeout 600 python bench.py --no-cpu-baseline --steps 60 > gpurun_out/bench_$v.log 2>&1; echo "bench $v: $(tail -1 gpurun_out/bench_$v.log | cut -c70-110)"
done
