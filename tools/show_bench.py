#!/usr/bin/env python
"""Print the interesting parts of bench.py JSON lines: python tools/show_bench.py file..."""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f):
        if line.startswith('{'):
            d = json.loads(line)
            print(f, 'N=%d' % d['n_gpus'], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'dev_ms', round(d.get('device_ms_per_step', 0), 4),
                  'e2e', round(d['e2e']['value'], 1), 'frac', round(d['roofline']['frac'], 4), 'kernel_ms', round(d['roofline']['kernel_ms'], 4))
            print('  phases', {k: round(v, 4) for k, v in d['phases_ms'].items()})
            if 'kernels' in d:
                print('  kernels', {k: (round(v['gbs'], 1), round(v['frac_of_hbm_peak'], 4)) for k, v in d['kernels'].items()})
            s = d.get('stmv')
            if s:
                print('  STMV value', round(s['value'], 2), 'ms', round(s['ms_per_step'], 3), 'dev', round(s['device_ms_per_step'], 3), 'e2e', round(s['e2e']['value'], 2),
                      'frac', round(s['roofline']['frac'], 4), s['roofline'].get('interacting_pairs_all_ranks', s['roofline'].get('interacting_pairs')))
                print('  phases', {k: round(v, 4) for k, v in s['phases_ms'].items()})
            if d.get('cpu_baseline'):
                print('  cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
        elif 'rror' in line or 'Trace' in line:
            print(line[:300])
