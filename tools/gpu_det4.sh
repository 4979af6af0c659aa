#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 -x -k "detect" > gpurun_out/pytest_det.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_det.log
grep -E "^E  |passed|failed|rc=|y2 detect|Error" gpurun_out/pytest_det.log | head -30
for c in 3 2 4; do echo "--- ctas/sm $c"; Y2_DETECT_CTAS_PER_SM=$c timeout 300 python tools/bench_detect.py 2>&1 | head -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['fused_us'], d['split_us'], d['split_frac_of_hbm_peak'])"; done
