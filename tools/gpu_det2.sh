#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:detect_ --csv --log-file gpurun_out/det_launches.csv python tools/bench_detect.py > gpurun_out/det_ncu.log 2>&1
echo "rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/det_launches.csv') if l.startswith('"'))]
h=rows[0]; ix={k:i for i,k in enumerate(h)}
d={}
for r in rows[1:]:
    d.setdefault((int(r[ix['ID']]), r[ix['Kernel Name']][:40]),{})[r[ix['Metric Name']]]=r[ix['Metric Value']]
for k in sorted(d)[:60]:
    print(k, d[k])
PY
