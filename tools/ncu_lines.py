"""Join an ncu SASS source page (per-instruction counts) with nvdisasm line info
-> executed warp-instructions and stall samples per CUDA source line.
  python tools/ncu_lines.py <rep> <kernel regex> <cubin> <mangled-substring> [top]
"""
import csv, io, re, subprocess, sys, collections

rep, kre, cubin, fun = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several kernels may follow each other: take the first block
hdr = rows[1]
ia, ii, isrc, ismp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
inst = []
for r in rows[2:]:
    if len(r) != len(hdr) or not r[ia].startswith('0x'):
        if inst:
            break
        continue
    inst.append((int(r[ia], 16), int(r[ii] or 0), int(r[ismp] or 0), r[isrc]))
base = inst[0][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
line_of = {}
cur_fun, cur_line = None, None
for l in dis.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m:
        cur_fun = m.group(1)
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split('/')[-1], int(m.group(2)))
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m and cur_fun and fun in cur_fun:
        line_of[int(m.group(1), 16)] = cur_line
agg = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for a, n, s, txt in inst:
    ln = line_of.get(a - base)
    agg[ln][0] += n
    agg[ln][1] += s
    tot_i += n
    tot_s += s
print(f'total warp-instructions {tot_i}, stall samples {tot_s}')
for ln, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{str(ln):28s} inst {n:11d} {n/tot_i:6.1%}   samples {s:7d} {s/max(tot_s,1):6.1%}')
