#!/usr/bin/env python3
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept
under profiles/ (the .ncu-rep files themselves are scratch).

    python tools/ncu_summary.py launches gpurun_out/launches_r01.csv > profiles/r01_launches.md
    python tools/ncu_summary.py full gpurun_out/prof_conv_fwd.ncu-rep > profiles/r01_conv_fwd.md
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_uniform.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
    'sm__cycles_active.avg', 'smsp__inst_executed.sum',
]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for row in csv.DictReader(lines):
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except (KeyError, ValueError):
            continue
        unit = row.get('Metric Unit', 'ns')
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3}.get(unit, v)
        name = row['Kernel Name'].split('(')[0].replace('void ', '')
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print('| kernel | launches | total us | share | avg us |')
    print('|---|---:|---:|---:|---:|')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print('| `{}` | {} | {:.1f} | {:.1f}% | {:.1f} |'.format(k[:70], cnt[k], v, 100 * v / total,
                                                                v / cnt[k]))
    print('\ntotal {:.1f} us over {} launches (ncu serialised, cold-cache: compare shares)'.format(
        total, sum(cnt.values())))


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('### `{}`\n'.format(r[hdr.index('Kernel Name')][:90]))
        print('| metric | value | unit |\n|---|---:|---|')
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print('| {} | {} | {} |'.format(k, r[i], units[i]))
        print()


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
