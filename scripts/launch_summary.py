#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r'\(.*', '', row['Kernel Name']); name = re.sub(r'^void rba::', '', name)
    v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else v * 1e3 if unit == 'ms' else v
    a = agg.setdefault(name, [0, 0.0, 1e9, 0]); a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':58s} {'launches':>8s} {'total_us':>10s} {'avg_us':>8s} {'min_us':>8s} {'max_us':>8s} {'share':>6s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:58]:58s} {a[0]:8d} {a[1]:10.1f} {a[1]/a[0]:8.1f} {a[2]:8.1f} {a[3]:8.1f} {100*a[1]/tot:5.1f}%")
