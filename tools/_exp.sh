timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_scripts_gpu.py -q -k "resize or fused_bn or imagenet or script" 2>&1 | grep -v "^$" | tail -30
