#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "conv1_u8" --timeout 300 > gpurun_out/pytest_conv1.log 2>&1; echo "conv1 test rc=$?"; tail -3 gpurun_out/pytest_conv1.log
timeout 100 python tools/bench_conv1.py
IS=608 timeout 100 python tools/bench_conv1.py
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_scripts_gpu.py -q --timeout 600 > gpurun_out/pytest_e2e.log 2>&1; echo "e2e rc=$?"; tail -5 gpurun_out/pytest_e2e.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
