#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/ (run here, no GPU needed).
  python tools/ncu_summary.py launches <launches.csv> <out.txt>
  python tools/ncu_summary.py full <report.ncu-rep> <out.txt>"""
import collections
import csv
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
           'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
           'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg',
           'sm__cycles_elapsed.max']


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    H = rows[hdr]
    ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0][:70]
        v = float(r[vi].replace(',', ''))
        v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, 'w') as f:
        f.write(f"# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({src})\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:72s} n={c:4d} total_us={t:11.1f} share={t / tot:6.1%} avg_us={t / c:9.1f}\n")
        f.write(f"TOTAL launches={sum(a[0] for a in agg.values())} total_us={tot:.1f}\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write(f"# ncu --set full --clock-control none ({src}); selected metrics per captured launch\n")
        for r in rows[2:]:
            f.write('---\n')
            f.write(f"{'Kernel Name':75s} {r[H.index('Kernel Name')]}\n")
            for m in METRICS:
                if m in H:
                    i = H.index(m)
                    f.write(f"{m:75s} {r[i]} {U[i]}\n")
    print(open(dst).read())


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
