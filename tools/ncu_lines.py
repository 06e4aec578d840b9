#!/usr/bin/env python
"""Aggregate an ncu source-page capture by CUDA source line.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-name-substring> [top-N]

ncu's CSV source page is per SASS instruction; nvdisasm -g gives the file:line of every SASS
instruction of the in-tree library (same build as profiled).  Joined by instruction order."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'diffvg_b200', 'libdiffvg_b200.so')


def sass_lines(kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(['cuobjdump', '-xelf', 'all', LIB], cwd=tmp, stdout=subprocess.DEVNULL)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith('.cubin') or '-' in f.split('.sm_')[0]:
            continue
        txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_kernel = None
        loc = None
        for line in txt.splitlines():
            m = re.match(r'\s*\.section\s+\.text\.(\S+?),', line)
            if m:
                cur_kernel = m.group(1)
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
            if m:
                if 'inlined at' in m.group(3) and loc is not None:
                    continue
                loc = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if cur_kernel and kernel_sub in cur_kernel and re.match(r'\s*/\*[0-9a-f]{4,}\*/', line):
                out.append((cur_kernel, loc, line.strip()))
    return out


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    col = sys.argv[4] if len(sys.argv) > 4 else '# Samples'   # e.g. stall_no_inst, stall_long_sb, stall_wait
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + ksub],
                         capture_output=True, text=True).stdout
    # the report may hold several kernels matching: split on the "Kernel Name" header rows
    blocks = re.split(r'(?m)^"Kernel Name",', txt)
    mangled = {'k_edge': 'k_edge', 'k_render<1>': 'k_renderILb1', 'k_render<0>': 'k_renderILb0',
               'k_wave_composite_px<1>': 'k_wave_composite_pxILb1', 'k_wave_composite_px<0>': 'k_wave_composite_pxILb0'}
    for blk in blocks[1:]:
        name = blk.splitlines()[0]
        rows = list(csv.reader(io.StringIO('\n'.join(blk.splitlines()[1:]))))
        hdr = rows[0]
        data = rows[1:]
        key = None
        norm = name.replace('(bool)', '').replace('true', '1').replace('false', '0')
        for k, v in mangled.items():
            if k in norm:
                key = v
        m2 = re.search(r'(k_\w+)<([^>]*)>', name)
        if key is None and m2:      # template<bool, ...>: the mangled name carries ILb0ELb1E...E
            bools = re.findall(r'\(bool\)([01])', m2.group(2))
            if bools:
                key = '%sI%sE' % (m2.group(1), ''.join('Lb%sE' % b for b in bools))
        sass = [s for s in sass_lines(key or ksub)]
        if len(sass) != len(data):
            print('warning: %d SASS rows in report vs %d in library for %s' % (len(data), len(sass), name[:60]))
        i_s = hdr.index(col)
        i_e = hdr.index('Instructions Executed')
        i_t = hdr.index('Thread Instructions Executed')
        agg = defaultdict(lambda: [0, 0, 0])
        tot = [0, 0, 0]
        for r, s in zip(data, sass):
            a = agg[s[1]]
            for j, i in enumerate((i_s, i_e, i_t)):
                v = int(r[i] or 0)
                a[j] += v
                tot[j] += v
        print('==', name[:90], 'samples', tot[0], 'warp-inst', tot[1], 'thread-inst', tot[2])
        for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            print('  %-22s %6.2f%% samples  %6.2f%% inst  thr/inst %.1f' % (
                '%s:%d' % loc if loc else '?', 100.0 * a[0] / max(tot[0], 1), 100.0 * a[1] / max(tot[1], 1), a[2] / max(a[1], 1)))


if __name__ == '__main__':
    main()
