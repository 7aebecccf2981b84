#!/usr/bin/env python
"""Per-kernel totals per step from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    n = r[ki].split('(')[0]
    agg[n][0] += 1
    agg[n][1] += float(r[vi].replace(',', '')) / 1e6
nb = max(1, agg[[k for k in agg if 'bow_score' in k][0]][0])
print('steps', nb)
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('%-50s %5.1f launches %8.3f ms' % (k[:50], v[0] / nb, v[1] / nb))
print('total ms/step %.3f' % (sum(v[1] for k, v in agg.items() if 'sample_table' not in k) / nb))
