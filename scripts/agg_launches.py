"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row['Kernel Name']
    val = float(row['Metric Value'].replace(',', ''))
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0}.get(row['Metric Unit'], 1e-6)
    a = agg.setdefault(name[:72], [0, 0.0])
    a[0] += 1
    a[1] += val * scale
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:74s} n={v[0]:4d} total={v[1]:9.3f} ms  avg={v[1] / v[0]:8.4f} ms  {100 * v[1] / tot:5.1f}%")
print("total", round(tot, 3), "ms")
