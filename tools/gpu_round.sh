#!/bin/bash
# Round-closing validation on one B200 (gpurun -- "Y2_HEAD=<sha> bash tools/gpu_round.sh <tag>"): the whole GPU test suite, smoke(),
# the driver's bench command, BASELINE configs[3], the reference arm, the training step (configs[4]), the decode+NMS
# microbench (configs[2]) and the ncu launch lists behind roofline.traffic.  Everything lands in gpurun_out/<tag>_*.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/${TAG}_bench.err | grep '^{' > gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --precision bf16x3 --single-mode --no-cpu-baseline 2>> gpurun_out/${TAG}_bench.err | grep '^{' > gpurun_out/${TAG}_bench_bf16x3.json
timeout 600 python bench.py --image-size 608 --batch 32 --steps 20 --warmup 5 --no-cpu-baseline 2>> gpurun_out/${TAG}_bench.err | grep '^{' > gpurun_out/${TAG}_bench_608.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 2>> gpurun_out/${TAG}_bench.err | grep '^{' > gpurun_out/${TAG}_bench_reference_arm.json
timeout 600 python bench.py --mode train --steps 10 --warmup 3 2>> gpurun_out/${TAG}_bench.err | grep '^{' > gpurun_out/${TAG}_bench_train_1gpu.json
timeout 300 python tools/bench_detect.py 2>> gpurun_out/${TAG}_bench.err | grep '^{' > gpurun_out/${TAG}_bench_detect.json
bash tools/profile_step.sh ${TAG} 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_train_raw.csv python bench.py --mode train --steps 2 --warmup 3 --no-graph --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_train_breakdown.py gpurun_out/${TAG}_train_raw.csv --detail > gpurun_out/${TAG}_train_breakdown.txt; rm -f gpurun_out/${TAG}_train_raw.csv
tail -c 400 gpurun_out/${TAG}_bench.err
python - <<PY
import json
for f in ('bench','bench_bf16x3','bench_608','bench_reference_arm','bench_train_1gpu'):
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json'%f).read().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), (d.get('roofline') or {}).get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
