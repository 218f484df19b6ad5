#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel family."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        name = re.sub(r'_GLOBAL__N__[0-9a-f_]+cu_[0-9a-f]+', '', name)[:100]
        val = float(row['Metric Value'].replace(',', ''))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{v[1] / 1e3:10.1f} us {v[0]:5d}x {100 * v[1] / tot:5.1f}%  avg {v[1] / v[0] / 1e3:8.1f} us  {k}')
    print(f'total {tot / 1e3:.1f} us over {sum(v[0] for v in agg.values())} launches')


if __name__ == '__main__':
    main(sys.argv[1])
