#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel table kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python profiles/ncu_summary.py raw.csv > profiles/rNN_xxx.md
"""
import csv
import sys

COLS = [
    ('gpu__time_duration.sum', 'time_us'),
    ('dram__bytes_read.sum', 'dram_rd_MB'),
    ('dram__bytes_write.sum', 'dram_wr_MB'),
    ('launch__registers_per_thread', 'regs'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ_%'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_%'),
    ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64_%'),
    ('lts__t_sector_hit_rate.pct', 'L2hit_%'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'st_long'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'st_short'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'st_bar'),
    ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'st_noinst'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'st_wait'),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print('| kernel | ' + ' | '.join(n for _, n in COLS) + ' | dram_GB/s |')
    print('|---|' + '---|' * (len(COLS) + 1))
    for r in data:
        name = r[idx['Kernel Name']].replace('<unnamed>::', '').replace('void ', '')
        name = name.replace('(int)', '').split('(')[0][:60]
        vals = []
        t = rd = wr = None
        for key, short in COLS:
            if key not in idx:
                vals.append('-')
                continue
            v = float(r[idx[key]].replace(',', ''))
            u = units[idx[key]]
            if short == 'time_us':
                v = v / 1000 if u in ('ns', 'nsecond') else v * 1000 if u in ('ms', 'msecond') else v
                t = v
            if short in ('dram_rd_MB', 'dram_wr_MB'):
                v = {'Gbyte': v * 1000, 'Kbyte': v / 1000, 'byte': v / 1e6}.get(u, v)
                if short == 'dram_rd_MB':
                    rd = v
                else:
                    wr = v
            vals.append(f'{v:.1f}' if abs(v) >= 10 else f'{v:.2f}')
        bw = (rd + wr) / t if t and rd is not None and wr is not None else 0   # MB/us = TB/s
        print(f'| {name} | ' + ' | '.join(vals) + f' | {bw * 1000:.0f} |')


if __name__ == '__main__':
    main(sys.argv[1])
