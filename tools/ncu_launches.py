"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
Usage: python tools/ncu_launches.py launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]
ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
agg = collections.OrderedDict()
tot = 0.0
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    name = re.sub(r'\(.*', '', r[ki])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print('total %.1f us over %d launches' % (tot, sum(a[0] for a in agg.values())))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-46s n=%3d total=%9.1f us avg=%8.1f share=%5.1f%%' % (k[:46], n, t, t / n, 100 * t / tot))
