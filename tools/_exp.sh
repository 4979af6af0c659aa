mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_training_gpu.py -q -x 2>&1 | grep -v "^$" | tail -6
for E in 1 0; do
if [ $E = 1 ]; then export Y2_WGRAD_NO_GROUP=1; else unset Y2_WGRAD_NO_GROUP; fi
timeout 300 python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_train_$E.json 2> gpurun_out/r2k_train_$E.err; tail -2 gpurun_out/r2k_train_$E.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2k_train_$E.json') if l.startswith('{')][-1])
print('no_group=$E train', round(d['value']), round(d['ms_per_step'],3), d['phases_ms'])
PY
done
unset Y2_WGRAD_NO_GROUP
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wgrad --csv --log-file gpurun_out/wg.csv python bench.py --mode train --steps 1 --warmup 3 --no-graph --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/wg.csv') if l.startswith('"'))]
ix={h:i for i,h in enumerate(rows[0])}
v=[float(r[ix['Metric Value']].replace(',',''))*{'ns':1e-3,'us':1,'ms':1e3}[r[ix['Metric Unit']]] for r in rows[1:]]
print('grouped per-launch us (L22..L1):', [round(t) for t in v[-22:]], 'sum', round(sum(v[-22:])))
PY
rm -f gpurun_out/wg.csv
