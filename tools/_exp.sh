mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py -q -x 2>&1 | grep -v "^$" | tail -6
timeout 300 python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_train.json 2> gpurun_out/r2j_train.err; tail -2 gpurun_out/r2j_train.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2j_train.json') if l.startswith('{')][-1])
print('train', round(d['value']), round(d['ms_per_step'],3), d['phases_ms'], 'e2e', round(d['e2e']['value']), d['launches_per_step'])
PY
