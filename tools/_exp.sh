mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_e2e_gpu.py -q -x -k "conv_split or small or pipelined" 2>&1 | tail -5
for G in 0 1; do
  if [ $G = 1 ]; then export Y2_CONV_STREAMK_X3_GENERIC=1; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --single-mode --precision bf16x3 --no-cpu-baseline --sustain-seconds 0 > gpurun_out/r2f_x3_g$G.json 2>gpurun_out/r2f_x3_g$G.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2f_x3_g$G.json') if l.startswith('{')][-1])
m=d['precision_modes']['bf16x3']
print('x3 generic_streamk=$G value %.0f ms %.4f e2e %.0f'%(d['value'], d['ms_per_step'], d['e2e']['value']), [round(x*1e3) for x in m['per_layer_ms']])
PY
done
unset Y2_CONV_STREAMK_X3_GENERIC
timeout 300 python bench.py --steps 20 --warmup 5 --single-mode --no-cpu-baseline --sustain-seconds 0 > gpurun_out/r2f_bf16.json 2>gpurun_out/r2f_bf16.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2f_bf16.json') if l.startswith('{')][-1])
print('bf16 value %.0f ms %.4f e2e %.0f'%(d['value'], d['ms_per_step'], d['e2e']['value']))
PY
