#!/bin/bash
# compute-sanitizer memcheck over the kernels added late in round 1 (pair kernels, split detection, fast BN passes)
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 700 -x -k "detect or streamk or affine_rows or bn_stats_fold or halo_pair" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|rc=" gpurun_out/sanitize_memcheck.log | head -20
