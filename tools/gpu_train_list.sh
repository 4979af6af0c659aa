#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 864 -c 288 --csv --log-file gpurun_out/launches_train.csv python tools/bench_train.py --steps 1 --warmup 3 > gpurun_out/train_under_ncu.log 2>&1
echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/launches_train.csv') if l.startswith('"'))]
h=rows[0]; ix={k:i for i,k in enumerate(h)}
agg=collections.OrderedDict()
tot=0
for r in rows[1:]:
    name=r[ix['Kernel Name']].replace('void ','').replace('y2::','')
    name=name.split('(')[0][:60]
    t=float(r[ix['Metric Value']])/1e3
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=t; tot+=t
print('total us', round(tot,1), 'launches', len(rows)-1)
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:24]:
    print('%-62s n=%3d  %8.1f us  %4.1f%%' % (k,n,t,100*t/tot))
PY
