"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch
count, total time and share.  Usage: python tools/summarize_launches.py launches.csv [skip_first_n]"""
import collections
import csv
import re
import sys


def main(path, skip=0, top=30):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        n += 1
        if n <= skip:
            continue
        name = row['Kernel Name']
        v = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1.0)
        short = re.sub(r'^void ', '', name)
        short = re.sub(r'\(.*', '', short)[:100]
        d = agg.setdefault(short, [0, 0.0])
        d[0] += 1
        d[1] += v
    tot = sum(v[1] for v in agg.values())
    print('%d launches (first %d skipped), total %.3f ms' % (n - skip, skip, tot / 1e6))
    print('%7s %11s %7s %9s  %s' % ('count', 'total ms', 'share', 'avg us', 'kernel'))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print('%7d %11.3f %6.1f%% %9.1f  %s' % (v[0], v[1] / 1e6, 100 * v[1] / tot, v[1] / v[0] / 1e3, k))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
