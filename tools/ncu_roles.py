#!/usr/bin/env python
"""Who waits on whom in a warp-specialised kernel: from an .ncu-rep (captured with --import-source on) list every
mbarrier wait (SYNCS...TRYWAIT) with its execution count -- executions beyond the number of waits are spins -- and
its stall samples, split by role region (producer = up to the last UTMALDG, epilogue = around LDTM, MMA = around
UTCHMMA), plus the headline metrics.   usage: ncu_roles.py report.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[hi]
isamp, iexec, isrc = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed'), hdr.index('Source')
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[isamp]), int(r[iexec]), r[isrc].strip()))
    except Exception:
        pass
tot = sum(d[0] for d in data)
def idx(pred):
    return [i for i, d in enumerate(data) if pred(d[2])]
tma, ldtm, mma = idx(lambda s: 'UTMALDG' in s), idx(lambda s: 'LDTM' in s), idx(lambda s: 'UTCHMMA' in s)
print('kernel', rows[hi - 1][1][:90] if hi else '', '| samples', tot)
print('regions: UTMALDG %s..%s  LDTM %s..%s  UTCHMMA %s..%s' % (tma[0] if tma else None, tma[-1] if tma else None,
      ldtm[0] if ldtm else None, ldtm[-1] if ldtm else None, mma[0] if mma else None, mma[-1] if mma else None))
for i, (s, e, t) in enumerate(data):
    if 'TRYWAIT' in t and e > 0:
        # samples attributed to the wait loop: the TRYWAIT and the following branch
        s2 = s + (data[i + 1][0] if i + 1 < len(data) else 0)
        role = 'producer' if tma and i < tma[-1] else ('epilogue' if ldtm and mma and i < mma[0] else 'mma')
        print('  #%5d %-9s exec=%9d samples=%6d (%4.1f%%)' % (i, role, e, s2, 100.0 * s2 / max(tot, 1)))
for name, lst in (('UTCHMMA', mma), ('UTMALDG', tma), ('LDTM', ldtm)):
    print('  %s executed: %d' % (name, sum(data[i][1] for i in lst)))
if mma:
    lo = max(i for i, d in enumerate(data) if 'TRYWAIT' in d[2] and i < mma[0] and d[1] > 0)
    # MMA-warp samples from the tmem_empty wait (two TRYWAITs before the first MMA) to the end of the loop
    tw = [i for i, d in enumerate(data) if 'TRYWAIT' in d[2] and i < mma[0] and d[1] > 0]
    start = tw[-2] if len(tw) >= 2 else lo
    end = mma[-1] + 60
    print('  MMA-warp region #%d..#%d samples=%d' % (start, end, sum(d[0] for d in data[start:end])))
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
d = dict(zip(rr[0], zip(rr[1], rr[2])))
for k in ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
          'smsp__inst_executed.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum', 'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed',
          'l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']:
    if k in d:
        print('  %-68s %s %s' % (k, d[k][1], d[k][0]))
