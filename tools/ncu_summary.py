#!/usr/bin/env python
"""Summarise an `ncu --set full` capture for profiles/:

    python tools/ncu_summary.py <report.ncu-rep> <out.txt> <out_traffic.json>

out.txt: one block per captured launch with the metrics the design discussion refers to; out_traffic.json: DRAM bytes
(read + write), duration and FP64-pipe utilisation of the LONGEST launch of each kernel (bench.py reads the newest
profiles/r*_ncu_traffic.json for `roofline.traffic` / `roofline.fp64_pipe_pct`)."""
import csv
import io
import json
import re
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
           'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
           'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
           'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']


def short(name):
    name = re.sub(r'^void ', '', name)
    name = name.replace('dvg::', '')
    m = re.match(r'(\w+)(<[^>]*>)?', name)
    base = m.group(1)
    t = m.group(2) or ''
    t = t.replace('(bool)1', 'true').replace('(bool)0', 'false').replace('(int)', '')
    if base.startswith('k_wave_'):      # bool template parameters print as 0 / 1 on the raw page; DVG_LAUNCH names them false / true
        t = t.replace('<1>', '<true>').replace('<0>', '<false>')
    return base + t


def main():
    rep, out_txt, out_json = sys.argv[1:4]
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--metrics', ','.join(METRICS)],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    i_name = hdr.index('Kernel Name')
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    longest = {}
    with open(out_txt, 'w') as f:
        f.write('ncu --set full --clock-control none, one fwd+bwd step of bench.py (painterly 2048 paths, 512^2, 4x4 spp); per launch, in launch order\n\n')
        for r in rows[2:]:
            f.write('== %s\n' % r[i_name])
            vals = {}
            for m, i in cols:
                vals[m] = r[i]
                f.write('  %-62s %s %s\n' % (m, r[i], units[i]))
            f.write('\n')
            unit_t = units[hdr.index('gpu__time_duration.sum')]
            t = float(vals['gpu__time_duration.sum'].replace(',', ''))
            ms = t / 1e6 if unit_t in ('ns', 'nsecond') else (t / 1e3 if unit_t in ('us', 'usecond') else t)

            def to_bytes(m):
                v = float(vals[m].replace(',', ''))
                u = units[hdr.index(m)]
                return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
            k = short(r[i_name])
            e = {'dram_bytes': to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum'), 'ms': ms,
                 'fp64_pipe_pct': float(vals['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'].replace(',', '')),
                 'lanes_per_instruction': float(vals['smsp__thread_inst_executed_per_inst_executed.ratio'].replace(',', ''))}
            if k not in longest or longest[k]['ms'] < ms:
                longest[k] = e
    json.dump({'source': 'ncu --set full capture summarised in %s: dram__bytes_read.sum + dram__bytes_write.sum, duration and FP64-pipe '
                         'utilisation of the LONGEST launch of each kernel in one step (for the exact-test kernel: the boundary-pass launch)' % out_txt,
               'kernels': longest}, open(out_json, 'w'), indent=1)
    for k, e in sorted(longest.items(), key=lambda kv: -kv[1]['ms']):
        print('%-34s %.3f ms  dram %.1f MB  fp64 pipe %.1f%%  lanes %.1f' % (k, e['ms'], e['dram_bytes'] / 1e6, e['fp64_pipe_pct'], e['lanes_per_instruction']))


if __name__ == '__main__':
    main()
