#!/usr/bin/env python
"""Summarise an ncu source-page CSV by code region (role) and list top stalled instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isamp, iexec, isrc = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed'), hdr.index('Source')
data = [(int(r[isamp]), int(r[iexec]), r[isrc].strip()) for r in rows[2:] if r[isamp].isdigit()]
tot = sum(d[0] for d in data)
print(rows[0][1][:80], 'total samples', tot)
# find role boundaries by marker instructions
def first(pred, start=0):
    for i in range(start, len(data)):
        if pred(data[i][2]): return i
    return len(data)
i_tma = first(lambda s: 'UTMALDG' in s)
i_ldtm = first(lambda s: 'LDTM' in s)
i_mma = first(lambda s: 'UTCHMMA' in s)
print('first UTMALDG @%d, LDTM @%d, UTCHMMA @%d' % (i_tma, i_ldtm, i_mma))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for s, ex, src in sorted(data, reverse=True)[:N]:
    idx = data.index((s, ex, src))
    print('%6d %5.1f%%  #%4d exec=%8d  %s' % (s, 100.0 * s / max(tot, 1), idx, ex, src[:90]))
