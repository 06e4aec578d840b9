#!/usr/bin/env python
"""SASS instruction count of a kernel by source line / file (code-size view; the render kernels are
instruction-fetch bound: ncu stall_no_inst ~27%).

    python tools/sass_size.py <kernel-substring> [top-N]"""
import sys
from collections import defaultdict
sys.path.insert(0, __import__('os').path.dirname(__file__))
import ncu_lines as nl

ksub = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sass = nl.sass_lines(ksub)
byfile = defaultdict(int)
byline = defaultdict(int)
for k, loc, _ in sass:
    byfile[loc[0] if loc else '?'] += 1
    byline[loc] += 1
print('total', len(sass), 'SASS instructions =', len(sass) * 16 // 1024, 'KB')
for f, n in sorted(byfile.items(), key=lambda kv: -kv[1]):
    print('  %-34s %6d' % (f, n))
print('top lines:')
for loc, n in sorted(byline.items(), key=lambda kv: -kv[1])[:top]:
    print('  %-34s %6d' % ('%s:%d' % loc if loc else '?', n))
