#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 200 -x -k "halo_pair" > gpurun_out/pytest_halo.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_halo.log
grep -E "^E  |passed|failed|rc=|y2 conv" gpurun_out/pytest_halo.log | head -20
bash tools/gpu_sanitize.sh
