#!/bin/bash
# Multi-GPU measurements on ONE box (run through `gpurun --gpus N -- bash tools/gpu_multi.sh N <tag>`): inference bench (both
# precision modes), training bench (configs[4]), decode+NMS microbench sharded strong / weak (configs[2]); N = 2 also runs the
# data-parallel correctness check.  Writes gpurun_out/<tag>_*_<N>gpu.json.
set -u
N=${1:-2}
TAG=${2:-r2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "1" ]; then TR="python"; fi
$TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_infer_${N}gpu.json 2> gpurun_out/${TAG}_infer_${N}gpu.err
tail -c 300 gpurun_out/${TAG}_infer_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${TAG}_bench_infer_${N}gpu.json').read().splitlines() if l.startswith("{")][-1])
print('infer N=$N', {k:(round(v['value']), round(v['e2e']), round(v['sustained_value'] or 0)) for k,v in d['precision_modes'].items()})
PY
$TR bench.py --mode train --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_train_${N}gpu.json 2> gpurun_out/${TAG}_train_${N}gpu.err
tail -c 300 gpurun_out/${TAG}_train_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/${TAG}_bench_train_${N}gpu.json').read().splitlines() if l.startswith("{")][-1])
print('train N=$N', round(d['value']), d['ms_per_step'], d['phases_ms'], d['allreduce'])
PY
for SH in strong weak; do
  $TR tools/bench_detect.py --shard $SH > gpurun_out/${TAG}_bench_detect_${SH}_${N}gpu.json 2> gpurun_out/${TAG}_detect_${N}gpu.err
  tail -c 300 gpurun_out/${TAG}_detect_${N}gpu.err
  head -1 gpurun_out/${TAG}_bench_detect_${SH}_${N}gpu.json | cut -c1-420
done
if [ "$N" = "2" ]; then
  $TR tools/ddp_check.py 2>&1 | tail -2 | tee gpurun_out/${TAG}_ddp_check_2gpu.txt
  python -m pytest tests/test_training_gpu.py -q -k ddp 2>&1 | tail -2
fi
