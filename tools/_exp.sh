mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py tests/test_kernels_gpu.py -q -x -k "wgrad or pack_weights or training_step or loss_decreases or cuda_graph" 2>&1 | grep -v "^$" | tail -12
timeout 300 python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_train.json 2> gpurun_out/r2i_train.err; tail -2 gpurun_out/r2i_train.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2i_train.json') if l.startswith('{')][-1])
print('train', round(d['value']), round(d['ms_per_step'],3), d['phases_ms'], 'e2e', round(d['e2e']['value']))
PY
