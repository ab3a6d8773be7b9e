"""Per kernel of an ncu report: IPC and the top warp-stall reasons (cycles a warp waits per issued
instruction).  python tools/ncu_stalls.py <file.ncu-rep>"""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
want = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d['Kernel Name'].split('(')[0]
    vals = sorted([(float((d[w] or '0').replace(',', '')), w[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')])
                   for w in want], reverse=True)[:5]
    ipc = d.get('smsp__inst_executed.avg.per_cycle_active', '?')
    dur = d.get('gpu__time_duration.sum', '?')
    print(f"{name:34s} {dur:>8s} us  ipc/smsp={ipc:>5s}  " + '  '.join(f'{n}={v:.1f}' for v, n in vals))
