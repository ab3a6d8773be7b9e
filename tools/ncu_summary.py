"""Summarise ncu outputs into small text files for profiles/.
  python tools/ncu_summary.py launches <csv>          -> per-kernel launch counts / device time shares
  python tools/ncu_summary.py rep <file.ncu-rep>      -> key metrics per captured kernel
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum', 'smsp__cycles_active.avg',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_op_read.sum',
        'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        agg[row['Kernel Name'].split('(')[0]].append(float(row['Metric Value'].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    print(f'# {path}: ncu --metrics gpu__time_duration.sum (cold-cache, serialised: compare shares)')
    print(f'{"kernel":44s} {"launches":>8s} {"total_us":>11s} {"mean_us":>9s} {"max_us":>9s} {"share":>7s}')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f'{k:44s} {len(v):8d} {sum(v)/1e3:11.1f} {sum(v)/len(v)/1e3:9.2f} {max(v)/1e3:9.2f} {sum(v)/tot:7.1%}')


def rep(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f'# {path}: ncu --set full --clock-control none')
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"\n== {d['Kernel Name'].split('(')[0]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                print(f'  {k:75s} {d[k]:>16s} {units[hdr.index(k)]}')


if __name__ == '__main__':
    {'launches': launches, 'rep': rep}[sys.argv[1]](sys.argv[2])
