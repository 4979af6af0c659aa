N=4; TAG=r2e
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_infer_${N}gpu.json 2> gpurun_out/${TAG}_infer_${N}gpu.err
timeout 200 $TR bench.py --mode train --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_train_${N}gpu.json 2> gpurun_out/${TAG}_train_${N}gpu.err
timeout 100 $TR tools/bench_detect.py --shard strong > gpurun_out/${TAG}_bench_detect_strong_${N}gpu.json 2> gpurun_out/${TAG}_detect_${N}gpu.err
timeout 100 $TR tools/bench_detect.py --shard weak > gpurun_out/${TAG}_bench_detect_weak_${N}gpu.json 2> gpurun_out/${TAG}_detect_${N}gpu.err
echo done
