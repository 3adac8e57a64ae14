"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).

    python tools/ncu_summary.py <tag>      # e.g. r01a
Reads gpurun_out/launches.csv (launch list of one step) and
gpurun_out/prof_hot.ncu-rep (--set full capture of the hot kernels)."""
import csv
import io
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles')
METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'l1tex__t_sector_hit_rate.pct',
    'lts__t_sector_hit_rate.pct',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def short(name):
    name = re.sub(r'\(.*', '', name)
    name = name.replace('void ', '').replace('<unnamed>::', '')
    return name.strip()


def launches(tag):
    path = os.path.join(ROOT, 'gpurun_out', 'launches.csv')
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(io.StringIO(''.join(lines))))
    hdr = rows[0]
    ni, vi, mi = hdr.index('Kernel Name'), hdr.index('Metric Value'), \
        hdr.index('Metric Name')
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[mi] != 'gpu__time_duration.sum':
            continue
        a = agg[short(r[ni])]
        a[0] += 1
        a[1] += float(r[vi].replace(',', '')) / 1e3      # ns -> us
    total = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, f'{tag}_launches_step.md'), 'w') as f:
        f.write(f'# {tag}: ncu launch list of ONE forward step (config 2, '
                'B=1)\n\n`ncu --profile-from-start off --metrics '
                'gpu__time_duration.sum --clock-control none python '
                'tools/profile_step.py --part step` -- per-launch times are '
                'cold-cache and serialised: compare SHARES.\n\n')
        f.write(f'total {sum(a[0] for a in agg.values())} launches, '
                f'{total / 1e3:.3f} ms\n\n| kernel | launches | us | share |\n'
                '|---|---:|---:|---:|\n')
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'| {k} | {n} | {us:.1f} | {100 * us / total:.1f}% |\n')
    print('wrote launches summary:', total / 1e3, 'ms')


def hot(tag):
    rep = os.path.join(ROOT, 'gpurun_out', 'prof_hot.ncu-rep')
    if not os.path.exists(rep):
        return
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(OUT, f'{tag}_hot_kernels.md'), 'w') as f:
        f.write(f'# {tag}: ncu --set full of the hot kernels on full-size '
                'tensors\n\n`ncu --profile-from-start off --set full '
                '--clock-control none --import-source on python '
                'tools/profile_step.py --part hot`\n\n')
        for r in rows[2:]:
            f.write(f"## {short(r[idx['Kernel Name']])}  grid "
                    f"{r[idx['Grid Size']]} block {r[idx['Block Size']]}\n\n")
            for m in METRICS:
                if m in idx:
                    f.write(f'- {m}: {r[idx[m]]} {units[idx[m]]}\n')
            f.write('\n')
    print('wrote hot kernel summary')


if __name__ == '__main__':
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    hot(tag)
