#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
`bench.py --no-graph` into the two files profiles/ keeps per round:

    python tools/ncu_launch_list.py gpurun_out/launches_infer.csv profiles/r1c

writes  <prefix>_launches_bench_step.csv   one bench step (flush memset .. next flush memset): id, kernel, us, DRAM bytes
        <prefix>_traffic.json              DRAM bytes of the conv launches of that step (bench.py's roofline.traffic) and
                                           the convs' share of the step under ncu
"""
import csv
import json
import re
import sys


def main():
    src, prefix = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = {}
    order = []
    for r in rows[1:]:
        k = int(r[ix['ID']])
        if k not in launches:
            name = r[ix['Kernel Name']]
            name = re.sub(r'^void ', '', name)
            depth = 0
            for i, ch in enumerate(name):          # cut the parameter list: first '(' outside template brackets
                depth += ch == '<'
                depth -= ch == '>'
                if ch == '(' and depth == 0:
                    name = name[:i]
                    break
            name = name.replace('y2::', '')
            launches[k] = dict(id=k, kernel=name)
            order.append(k)
        v = float(r[ix['Metric Value']].replace(',', ''))
        m, u = r[ix['Metric Name']], r[ix['Metric Unit']]
        if m == 'gpu__time_duration.sum':
            launches[k]['us'] = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[u]
        else:
            launches[k][m] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    seq = [launches[k] for k in order]
    # one step = from a flush memset (FillFunctor<unsigned char>) to the next one, taking the LAST complete step
    flushes = [i for i, l in enumerate(seq) if 'FillFunctor<unsigned char>' in l['kernel']]
    assert len(flushes) >= 2, 'need two L2-flush launches to delimit a step'
    a, b = flushes[-2], flushes[-1]
    step = seq[a:b + 1]
    with open(prefix + '_launches_bench_step.csv', 'w') as f:
        f.write('id,kernel,grid_time_us,dram_read_bytes,dram_write_bytes\n')
        for l in step:
            f.write('%d,"%s",%.3f,%d,%d\n' % (l['id'], l['kernel'], l['us'], l.get('dram__bytes_read.sum', 0),
                                              l.get('dram__bytes_write.sum', 0)))
    body = step[1:-1]
    convs = [l for l in body if l['kernel'].startswith('conv')]
    t_all = sum(l['us'] for l in body)
    t_conv = sum(l['us'] for l in convs)
    out = dict(conv_dram_bytes_per_step=sum(l.get('dram__bytes_read.sum', 0) + l.get('dram__bytes_write.sum', 0) for l in convs),
               conv_launches_per_step=len(convs), launches_per_step=len(body), step_us_under_ncu=t_all,
               conv_share_of_step_under_ncu=t_conv / t_all,
               source='%s_launches_bench_step.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,'
                      'dram__bytes_write.sum --clock-control none, python bench.py --steps 3 --warmup 3 --no-graph)' % prefix)
    json.dump(out, open(prefix + '_traffic.json', 'w'), indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
