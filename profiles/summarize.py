#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed here.

    python profiles/summarize.py launches gpurun_out/<name>_launches.csv
    python profiles/summarize.py full     gpurun_out/<name>.ncu-rep

`launches`: per-kernel launch count, total and share of device time (from the
`--metrics gpu__time_duration.sum` pass).  `full`: the roofline-relevant counters of each
captured launch plus the executed-instruction mix by SASS opcode (source page).
"""
import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'smsp__inst_executed.sum',
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[mi].replace(',', ''))
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1., 's': 1e3}.get(r[ui], 1.)
        name = r[ki].split('(')[0].replace('void ', '').replace('unnamed>::', '')
        agg[name][0] += 1
        agg[name][1] += v
    total = sum(v[1] for v in agg.values())
    print('| kernel | launches | total ms | share | avg ms |')
    print('|---|---:|---:|---:|---:|')
    for name, (count, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `{}` | {} | {:.2f} | {:.1f}% | {:.3f} |'.format(name, count, ms, 100 * ms / total,
                                                               ms / count))


def full(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for idx, r in enumerate(rows[2:]):
        print('### launch {}: `{}`'.format(idx, r[hdr.index('Kernel Name')][:110]))
        print()
        for k in RAW_KEYS:
            if k in hdr:
                i = hdr.index(k)
                print('- {} = {} {}'.format(k, r[i], units[i]))
        print()
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    mix, total, k = collections.Counter(), 0, 0
    for r in csv.reader(io.StringIO(src)):
        if r and r[0] == 'Kernel Name':
            k += 1
            continue
        if k != 1 or len(r) < 6 or r[0] == 'Address':
            continue
        parts = r[1].split()
        op = parts[1] if parts[0].startswith('@') else parts[0]
        op = op.split('.')[0] + ('.MOV' if '.MOV' in op else '')
        n = int(r[5])
        mix[op] += n
        total += n
    print('### executed warp instructions by opcode (launch 0), total {}'.format(total))
    print()
    print('| opcode | warp instructions | share |')
    print('|---|---:|---:|')
    for op, n in mix.most_common(16):
        print('| {} | {} | {:.1f}% |'.format(op, n, 100. * n / max(1, total)))


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
