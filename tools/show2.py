#!/usr/bin/env python
"""Headline numbers of a bench.py JSON line (main workload + sub-objects): python tools/show2.py file.json ..."""
import json, sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'N=%d' % d['n_gpus'], d['config']['workload'], 'value', round(d['value'], 1), 'dev_ms', round(d['device_ms_per_step'], 4), 'e2e', round(d['e2e']['value'], 1),
          'frac', round(d['roofline']['frac'], 4), 'kernel_ms', round(d['roofline']['kernel_ms'], 4), d['roofline']['kernel'][:15])
    if d.get('phases_ms'):
        print('   ', d['phases_ms'])
    for k in ('stmv', 'apoa1', 'fep', 'replicas_dhfr'):
        s = d.get(k)
        if s:
            print('  ', k, 'value', round(s['value'], 1), 'dev', round(s.get('device_ms_per_step', 0), 4), 'e2e', round(s['e2e']['value'], 1), 'frac', round(s['roofline']['frac'], 4), s.get('phases_ms') or '')
