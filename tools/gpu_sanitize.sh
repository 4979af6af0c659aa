#!/bin/bash
# compute-sanitizer memcheck over the kernels added in round 2 (small shapes; the big-map cases are left out: memcheck is ~50x
# slower).  gpurun -- "bash tools/gpu_sanitize.sh r2" -> gpurun_out/<tag>_compute_sanitizer_memcheck.txt
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x \
  "tests/test_parity_gpu.py::test_conv_split_vs_oracle" "tests/test_parity_gpu.py::test_conv1_u8_pool_split_vs_oracle" \
  "tests/test_parity_gpu.py::test_affine_split_output" "tests/test_training_gpu.py::test_wgrad_first_layer" \
  "tests/test_training_gpu.py::test_dgrad_and_wgrad_vs_autograd" "tests/test_training_gpu.py::test_wgrad_split_k_large_map" \
  "tests/test_kernels_gpu.py::test_resize_bilinear_u8_bit_exact_vs_cv2" "tests/test_kernels_gpu.py::test_conv_streamk_fused_bn_statistics" \
  "tests/test_kernels_gpu.py::test_pack_weights_tiled_equals_generic" "tests/test_kernels_gpu.py::test_adam_step_ex_scale_zero_and_device_lr" \
  -k "not 416 and not 1024-1024" > gpurun_out/${TAG}_compute_sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/${TAG}_compute_sanitizer_memcheck.txt
tail -5 gpurun_out/${TAG}_compute_sanitizer_memcheck.txt
