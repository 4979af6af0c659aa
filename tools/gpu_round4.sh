#!/bin/bash
# round-1 closing run: GPU tests, smoke, both bench arms, 608 config, microbenches, training bench, launch list with DRAM
# bytes, full captures of the pair kernels (stream-K L19, halo L2 / L4) and the detection kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|rc=" gpurun_out/pytest_gpu.log | head -30
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 600 python bench.py --image-size 608 --batch 32 --no-cpu-baseline > gpurun_out/bench_608.log 2>&1; echo "bench 608 rc=$?"; tail -1 gpurun_out/bench_608.log | cut -c1-1500
timeout 300 python tools/bench_detect.py > gpurun_out/bench_detect.log 2>&1; echo "detect rc=$?"; tail -2 gpurun_out/bench_detect.log
timeout 300 python tools/run_layer.py L1 L2 L3 L4 L5 L6 L7 L8 L9 L10 L13 L14 L15 L19 L22 --iters 20 > gpurun_out/layers.log 2>&1; cat gpurun_out/layers.log
timeout 600 python tools/bench_train.py > gpurun_out/bench_train.log 2>&1; echo "train rc=$?"; tail -1 gpurun_out/bench_train.log | cut -c1-1200
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 90 -c 120 --csv --log-file gpurun_out/launches_infer.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
for L in L19 L2 L4; do
  EXTRA=""; [ $L = L19 ] && EXTRA="--raw"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv -s 1 -c 1 -f -o gpurun_out/ncu_full_$L python tools/run_layer.py $L --iters 1 $EXTRA > gpurun_out/ncu_full_$L.log 2>&1; echo "ncu $L rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:detect_decode -s 2 -c 1 -f -o gpurun_out/ncu_full_detect_decode python tools/bench_detect.py > gpurun_out/ncu_full_dd.log 2>&1; echo "ncu decode rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:detect_nms -s 2 -c 1 -f -o gpurun_out/ncu_full_detect_nms python tools/bench_detect.py > gpurun_out/ncu_full_dn.log 2>&1; echo "ncu nms rc=$?"
