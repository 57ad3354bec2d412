"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel. Usage: python tools/launch_shares.py file.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r'\(.*', '', r[kn]).replace('void <unnamed>::', '').replace('<unnamed>::', '')
    v = float(r[mv].replace(',', ''))
    v = v / 1e3 if r[mu] == 'ns' else (v * 1e3 if r[mu] == 'ms' else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-58s n=%3d total %8.1f us  avg %7.1f us  share %.3f' % (k[:58], n, t, t / n, t / tot))
print('total %.1f us over %d launches' % (tot, sum(n for n, _ in agg.values())))
