"""Aggregate tools/ncu_lines.py output (stdin) by named source-line ranges of one file.
  python tools/ncu_lines.py <rep> <kernel> <cubin> <mangled> 1000 | python tools/ncu_regions.py raster.cu 860:prologue 886:empty ...
Each argument START:NAME opens a region that ends where the next one starts."""
import sys, re, collections
fname = sys.argv[1]
marks = sorted((int(a.split(':')[0]), a.split(':')[1]) for a in sys.argv[2:])
reg = collections.Counter(); smp = collections.Counter(); tot = 0; tots = 0
for l in sys.stdin:
    m = re.match(r"\('([^']+)', (\d+)\)\s*inst\s+(\d+)\s+[\d.]+%\s+samples\s+(\d+)", l)
    if not m:
        m0 = re.match(r"None\s+inst\s+(\d+)\s+[\d.]+%\s+samples\s+(\d+)", l)
        if m0:
            reg['(no line)'] += int(m0.group(1)); smp['(no line)'] += int(m0.group(2)); tot += int(m0.group(1)); tots += int(m0.group(2))
        continue
    fn, ln, n, s = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4))
    tot += n; tots += s
    name = fn
    if fn == fname:
        name = 'before'
        for st, nm in marks:
            if ln >= st:
                name = nm
    reg[name] += n; smp[name] += s
for k, v in reg.most_common():
    print(f'{k:32s} inst {v:10d} {100*v/tot:5.1f}%   samples {smp[k]:6d} {100*smp[k]/max(tots,1):5.1f}%')
