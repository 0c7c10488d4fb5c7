#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel family (template arguments dropped) and, with
--detail, per (kernel, grid) in launch order.

    python tools/launch_summary.py gpurun_out/r02_train_launches.csv [--detail] > profiles/<name>.md
"""
import collections
import csv
import re
import sys


def rows(path):
    with open(path, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        value = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        us = value / 1e3 if unit in ('ns', 'nsecond') else value * (1e3 if unit in ('ms', 'msecond') else 1.0)
        yield r['Kernel Name'], r.get('Grid Size', ''), us


def family(name):
    name = re.sub(r'^void\s+', '', name)
    name = re.sub(r'\(.*$', '', name)
    name = re.sub(r'^(yb::|at::native::|\(anonymous namespace\)::)+', '', name)
    return re.sub(r'<.*$', '', name)


def main():
    path = sys.argv[1]
    detail = '--detail' in sys.argv
    data = list(rows(path))
    total = sum(us for _, _, us in data)
    fam = collections.OrderedDict()
    for name, grid, us in data:
        e = fam.setdefault(family(name), [0, 0.0])
        e[0] += 1
        e[1] += us
    print('Launch list `%s`: %d launches, %.1f us (cold-cache, serialised under ncu: compare shares).\n' % (path.split('/')[-1], len(data), total))
    print('| kernel | launches | total us | share |\n|---|---|---|---|')
    for k, (n, us) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print('| %s | %d | %.1f | %.3f |' % (k, n, us, us / total))
    if detail:
        print('\nIn launch order:\n\n| # | kernel | grid | us |\n|---|---|---|---|')
        for i, (name, grid, us) in enumerate(data):
            short = re.sub(r'\(.*$', '', re.sub(r'^void\s+', '', name))
            print('| %d | %s | %s | %.1f |' % (i, short, grid, us))


if __name__ == '__main__':
    main()
