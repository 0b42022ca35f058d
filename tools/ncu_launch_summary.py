#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): launches, summed and mean duration, share."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors='replace')))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hi]; ik = h.index('Kernel Name'); iv = h.index('Metric Value'); iu = h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= iv: continue
    v = float(r[iv].replace(',', ''))
    if r[iu] == 'ns': v /= 1e3
    elif r[iu] == 'ms': v *= 1e3
    name = r[ik].split('(')[0].replace('<unnamed>::', '')[-48:]
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print(f'{k:50s} n={len(v):3d} sum={sum(v):9.1f} us  mean={sum(v)/len(v):8.1f} us  share={sum(v)/tot*100:5.1f}%')
print(f'total {tot:.1f} us')
