#!/usr/bin/env python3
"""tools/launch_summary.py launches.csv -- per-kernel totals from `ncu --metrics gpu__time_duration.sum --csv`."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    name = row['Kernel Name'].split('(')[0]
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else v * 1e3 if unit == 'ms' else v
    agg[name].append(v)
tot = sum(sum(v) for v in agg.values())
print("%-44s %5s %10s %9s %9s %7s" % ("kernel", "n", "avg us", "min us", "max us", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-44s %5d %10.2f %9.2f %9.2f %6.1f%%" % (k[:44], len(v), sum(v) / len(v), min(v), max(v), 100 * sum(v) / tot))
