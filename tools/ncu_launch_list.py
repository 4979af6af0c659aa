#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
`bench.py --no-graph` into the two files profiles/ keeps per round:

    python tools/ncu_launch_list.py gpurun_out/launches_infer.csv profiles/r2a [bf16|bf16x3]

writes  <prefix>_launches_bench_step_<precision>.csv   one bench step (flush memset .. next flush memset): id, kernel, us, DRAM bytes
        <prefix>_traffic_<precision>.json              DRAM bytes of the conv launches of that step (bench.py's
                                                       roofline.traffic), the convs' share of the step under ncu, and the
                                                       git HEAD the library was built from (env Y2_HEAD, set by
                                                       tools/profile_step.sh)
"""
import csv
import json
import os
import re
import sys


def main():
    src, prefix = sys.argv[1], sys.argv[2]
    precision = sys.argv[3] if len(sys.argv) > 3 else 'bf16'
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
    # intervals between consecutive flushes; some hold several steps (bench.py's e2e loop does not flush): take the LAST
    # interval of the most common length, i.e. exactly one step
    spans = [(flushes[i], flushes[i + 1]) for i in range(len(flushes) - 1) if flushes[i + 1] - flushes[i] > 2]
    lengths = [y - x for x, y in spans]
    mode = max(set(lengths), key=lambda v: (lengths.count(v), -v))
    a, b = [sp for sp in spans if sp[1] - sp[0] == mode][-1]
    step = seq[a:b + 1]
    with open(prefix + '_launches_bench_step_%s.csv' % precision, 'w') as f:
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
               conv_share_of_step_under_ncu=t_conv / t_all, precision=precision, head=os.environ.get('Y2_HEAD'),
               source='%s_launches_bench_step_%s.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,'
                      'dram__bytes_write.sum --clock-control none, python bench.py --steps 3 --warmup 3 --no-graph '
                      '--single-mode --precision %s)' % (os.path.basename(prefix), precision, precision))
    json.dump(out, open(prefix + '_traffic_%s.json' % precision, 'w'), indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
