#!/usr/bin/env python
"""Aggregates an ncu --csv launch list (gpu__time_duration + inst metrics) per kernel.
usage: tools/wave_breakdown.py gpurun_out/launches_x.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
data = collections.defaultdict(dict)
for r in rows[1:]:
    try:
        data[(int(r[ii]), r[ki].split('(')[0])][r[mi]] = float(r[vi].replace(',', ''))
    except ValueError:
        pass
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for (_i, k), m in sorted(data.items()):
    for mn, v in m.items():
        agg[k][mn].append(v)
tot = sum(sum(m['gpu__time_duration.sum']) for m in agg.values())
for k, m in agg.items():
    t = m['gpu__time_duration.sum']
    n = len(t)
    g = lambda name: sum(m.get(name, [0])) / max(1, len(m.get(name, [0])))
    print(f"{k:28} n={n:4} avg_us={sum(t)/n/1e3:8.1f} share={100*sum(t)/tot:5.1f}% inst/launch={g('smsp__inst_executed.sum')/1e6:7.2f}M "
          f"lanes={g('smsp__thread_inst_executed_per_inst_executed.ratio'):5.1f} issue%={g('smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} "
          f"dram_MB={(g('dram__bytes_read.sum')+g('dram__bytes_write.sum'))/1e6:7.1f}")
