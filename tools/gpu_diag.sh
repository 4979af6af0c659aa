#!/bin/bash
# bring-up run on the GPU box: kernel parity tests + per-case conv diagnostics (one process per
# case so a trapped kernel cannot poison the next case).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 300 > gpurun_out/pytest_kernels.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_kernels.log
tail -5 gpurun_out/pytest_kernels.log
: > gpurun_out/diag_conv.log
for c in "$@"; do
  timeout 120 python tools/diag_conv.py $c >> gpurun_out/diag_conv.log 2>&1
  echo "case $c rc=$?" >> gpurun_out/diag_conv.log
done
grep -E '"case"|rc=' gpurun_out/diag_conv.log | cut -c1-400
