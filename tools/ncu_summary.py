#!/usr/bin/env python
"""Key metrics of an .ncu-rep (raw page) as a small text summary for profiles/.
usage: ncu_summary.py report.ncu-rep [algorithmic_flops] [algorithmic_bytes]"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, zip(units, r)))
    print('kernel:', d['Kernel Name'][1])
    for k in KEYS:
        if k in d:
            print('  %-70s %s %s' % (k, d[k][1], d[k][0]))
    t_us = float(d['gpu__time_duration.sum'][1])
    if d['gpu__time_duration.sum'][0] == 'ms': t_us *= 1e3
    if d['gpu__time_duration.sum'][0] == 'ns': t_us *= 1e-3
    def tobytes(k):
        u, v = d[k]; v = float(v)
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    traffic = tobytes('dram__bytes_read.sum') + tobytes('dram__bytes_write.sum')
    print('  traffic (dram read+write) per launch: %.3f MB -> %.1f GB/s under ncu' % (traffic / 1e6, traffic / t_us / 1e3))
    if len(sys.argv) > 2 and float(sys.argv[2]) > 0:
        print('  algorithmic %.3f GFLOP -> %.1f TFLOP/s under ncu (cold, serialised)' % (float(sys.argv[2]) / 1e9, float(sys.argv[2]) / t_us / 1e6))
    if len(sys.argv) > 3:
        print('  algorithmic bytes %.3f MB (traffic/algorithmic = %.2f)' % (float(sys.argv[3]) / 1e6, traffic / float(sys.argv[3])))
