#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time and launches per kernel name."""
import collections
import csv
import sys


def summarise(fn, top=30):
    with open(fn) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[row['Metric Unit']]
        a = agg.setdefault(row['Kernel Name'][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(t for _, t in agg.values())
    print(f'{fn}: {sum(c for c, _ in agg.values())} launches, {tot:.3f} ms')
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f'{t:10.3f} ms {100 * t / tot:5.1f}% {c:6d}  {k}')


if __name__ == '__main__':
    for fn in sys.argv[1:]:
        summarise(fn)
