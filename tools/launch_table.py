#!/usr/bin/env python
"""Per-kernel mean duration from an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_table.py file.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
data = []
for r in rows:
    if r and r[0] == 'ID':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for d in data[len(data)//2:]:
    n = d['Kernel Name'][:64]
    v = float(d['Metric Value'].replace(',', ''))
    u = d['Metric Unit']
    v = v/1000 if u == 'ns' else v*1000 if u == 'ms' else v
    agg.setdefault((n, d['Grid Size'], d['Block Size']), []).append(v)
for (n, g, b), v in agg.items():
    print('%-66s %-18s %-14s n=%3d mean %9.2f us min %9.2f' % (n, g, b, len(v), sum(v)/len(v), min(v)))
