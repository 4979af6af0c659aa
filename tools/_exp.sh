mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_training_gpu.py tests/test_scripts_gpu.py tests/test_parity_gpu.py tests/test_kernels_gpu.py tests/test_e2e_gpu.py -q -x -k "not detections_baseline" 2>&1 | tail -12
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2d_train_graph.json 2> gpurun_out/r2d_train_graph.err; tail -3 gpurun_out/r2d_train_graph.err; cat gpurun_out/r2d_train_graph.json | cut -c1-2500
timeout 600 python bench.py --mode train --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r2d_train_eager.json 2> gpurun_out/r2d_train_eager.err; tail -3 gpurun_out/r2d_train_eager.err; cat gpurun_out/r2d_train_eager.json | cut -c1-1200
