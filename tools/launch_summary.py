#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`):

    python tools/launch_summary.py <launches.csv> <out.txt> ["<the command that was profiled>"]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    src, out = sys.argv[1], sys.argv[2]
    cmd = sys.argv[3] if len(sys.argv) > 3 else 'python bench.py --steps 2 --warmup 3 --no-cpu-baseline'
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    tot = OrderedDict()
    for r in rows:
        name = re.sub(r'\(.*', '', r[kn].replace('dvg::', '').replace('void ', '')) if r[kn].startswith(('dvg::', 'void dvg::')) \
            else re.sub(r'\(.*', '', r[kn])[:100]
        name = name.replace('(bool)', '').replace('(int)', '')
        c, t = tot.get(name, (0, 0.0))
        tot[name] = (c + 1, t + float(r[mv].replace(',', '')) / 1e3)
    total = sum(t for _, t in tot.values())
    with open(out, 'w') as f:
        f.write('ncu --metrics gpu__time_duration.sum --clock-control none -c %d %s: first %d launches (serialised, cold cache: '
                'shares, not absolutes)\n' % (len(rows), cmd, len(rows)))
        for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write('%-60s launches %4d  total %10.1f us  share %5.1f%%\n' % (name, c, t, 100 * t / total))


if __name__ == '__main__':
    main()
