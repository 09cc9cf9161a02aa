#!/usr/bin/env python
"""Turn the raw ncu output of a gpurun call (gpurun_out/, scratch) into the small summaries kept under profiles/.

    python tools/summarise_profiles.py --launches gpurun_out/r01_loss_launches.csv --out profiles/r01_loss
    python tools/summarise_profiles.py --full gpurun_out/r01_final.ncu-rep --out profiles/r01_ncu_full_loss_assign.csv

--launches: the CSV log of `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...`
            -> <out>_launches.csv (copy) + <out>_launch_summary.csv (kernel, launches, avg_us, total_us, share)
--full    : an `ncu --set full` report (or its `--page raw --csv` export) -> key metrics per kernel as CSV, and profiles/traffic.json
            (dram__bytes_read.sum + dram__bytes_write.sum per launch; bench.py reads it for roofline.traffic)
"""
import argparse
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        # warp-state sampling: average warps stalled per issue slot, by reason (ncu --set full, WarpStateStats)
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio']
TO_BYTES = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def csv_rows(text):
    lines = [ln for ln in text.splitlines() if ln.startswith('"')]
    return list(csv.reader(io.StringIO('\n'.join(lines))))


def launches(path, out):
    rows = csv_rows(open(path).read())
    hdr = rows[0]
    k, v = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = {}
    for r in rows[1:]:
        if len(r) <= v:
            continue
        name = r[k][:70]
        us = float(r[v].replace(',', ''))
        unit = r[hdr.index('Metric Unit')]
        us *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1.0)
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + us)
    ours = sum(t for name, (n, t) in agg.items() if 'cldet::' in name)
    if os.path.abspath(path) != os.path.abspath(out + '_launches.csv'):
        shutil.copyfile(path, out + '_launches.csv')
    with open(out + '_launch_summary.csv', 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'launches', 'avg_us', 'total_us', 'share_of_cldet_kernels'])
        for name, (n, t) in agg.items():
            w.writerow([name, n, '%.2f' % (t / n), '%.2f' % t, '%.3f' % (t / ours) if 'cldet::' in name and ours else ''])
    print('wrote', out + '_launch_summary.csv')


def full(path, out):
    if path.endswith('.csv'):          # the raw page exported on the GPU box (the .ncu-rep with sources is ~32 MB per kernel)
        txt = open(path).read()
    else:
        txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = csv_rows(txt)
    hdr, units = rows[0], rows[1]
    kn = hdr.index('Kernel Name')
    kernels = rows[2:]
    traffic = {}
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit'] + [r[kn][:60] for r in kernels])
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                w.writerow([key, units[i]] + [r[i] for r in kernels])
    ir, iw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    for r in kernels:
        short = r[kn].split('(')[0].split('<')[0].split()[-1].replace('cldet::', '')
        traffic[short] = float(r[ir]) * TO_BYTES[units[ir]] + float(r[iw]) * TO_BYTES[units[iw]]
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    old = json.load(open(tp)) if os.path.exists(tp) else {}
    old.update(traffic)
    json.dump(old, open(tp, 'w'), indent=1)
    print('wrote', out, 'and', tp, traffic)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--launches')
    ap.add_argument('--full')
    ap.add_argument('--out', required=True)
    a = ap.parse_args()
    if a.launches:
        launches(a.launches, a.out)
    if a.full:
        full(a.full, a.out)
    return 0


if __name__ == '__main__':
    sys.exit(main())
