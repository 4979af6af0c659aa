#!/usr/bin/env python
"""Top sampled SASS instructions from an `ncu --page source --csv` dump (stdin or file)."""
import csv, sys
f = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
rows = list(csv.reader(f))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index('Address'), hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
data = []
for idx, r in enumerate(rows[2:]):
    try:
        data.append((int(r[isamp]), idx, r[isrc].strip(), int(r[iexec])))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print('total samples', tot)
for s, idx, src, ex in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print('%6d %5.1f%%  #%4d exec=%8d  %s' % (s, 100.0 * s / max(tot, 1), idx, ex, src[:100]))
