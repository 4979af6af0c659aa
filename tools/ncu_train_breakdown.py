#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --mode train --no-graph` by kernel for
ONE training step (the launches between the last two L2-flush memsets that enclose the most common launch count).
    python tools/ncu_train_breakdown.py gpurun_out/train_launches.csv [--detail] > profiles/r2_train_breakdown.txt
--detail appends the launches of the contraction kernels (forward conv, data gradient, weight gradient) in launch order."""
import csv
import re
import sys
from collections import OrderedDict


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    ix = {h: i for i, h in enumerate(rows[0])}
    seq = []
    for r in rows[1:]:
        if r[ix['Metric Name']] != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'^void ', '', r[ix['Kernel Name']])
        depth = 0
        for i, ch in enumerate(name):
            depth += ch == '<'
            depth -= ch == '>'
            if ch == '(' and depth == 0:
                name = name[:i]
                break
        v = float(r[ix['Metric Value']].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}[r[ix['Metric Unit']]]
        seq.append((name.replace('y2::', ''), v))
    fl = [i for i, (n, _) in enumerate(seq) if 'FillFunctor<unsigned char>' in n]
    spans = [(fl[i], fl[i + 1]) for i in range(len(fl) - 1) if fl[i + 1] - fl[i] > 50]
    lengths = [b - a for a, b in spans]
    mode = max(set(lengths), key=lengths.count)
    a, b = [sp for sp in spans if sp[1] - sp[0] == mode][-1]
    step = seq[a + 1:b]
    agg = OrderedDict()
    for n, v in step:
        k = re.sub(r'<.*', '', n)
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + v)
    total = sum(v for _, v in step)
    print('one training step under ncu (cold-cache, serialised launches): %d launches, %.1f us' % (len(step), total))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('  %-44s x%-3d %9.1f us  %5.1f %%' % (k, c, t, 100 * t / total))
    if '--detail' in sys.argv:
        print('contraction kernels in launch order (forward layers 1..22, then backward 22..1: wgrad, dgrad):')
        for i, (n, v) in enumerate(step):
            if n.startswith('conv'):
                print('  #%-3d %-60s %8.1f us' % (i, n[:60], v))


if __name__ == '__main__':
    main()
