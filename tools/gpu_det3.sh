#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 -x -k "detect" > gpurun_out/pytest_det.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_det.log
grep -E "^E  |passed|failed|rc=|y2 detect|Error" gpurun_out/pytest_det.log | head -30
timeout 300 python tools/bench_detect.py 2>&1 | tee gpurun_out/bench_detect.log | tail -3 | cut -c1-420
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:detect_ -s 46 -c 6 --csv --log-file gpurun_out/det_launches.csv python tools/bench_detect.py > gpurun_out/det_ncu.log 2>&1
grep -E "gpu__time|inst_executed" gpurun_out/det_launches.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-120
