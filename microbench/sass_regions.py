#!/usr/bin/env python
"""Dynamic instructions per region of a kernel's SASS (chunks of 100 instructions in address order),
from an ncu source-page CSV joined with nvdisasm -g line info of the same build.
  sass_regions.py src.csv dis.txt <mangled kernel name prefix> <probes>"""
import csv, re, sys
src_csv, dis_txt, kname, probes = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
lines = open(dis_txt).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('//--------------------- .text.' + kname))
end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith('//--------------------- '))
cur, seq = None, []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        if 'inlined' not in m.group(3):
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);', l)
    if m:
        seq.append((m.group(2).strip(), cur))
assert len(seq) <= len(data), (len(seq), len(data))
tot = sum(int(r[ix['Instructions Executed']]) for r in data)
print(f"{len(seq)} SASS instructions, {tot / probes:.2f} warp instr / probe")
chunk = 100
for k in range(0, len(seq), chunk):
    rng = range(k, min(k + chunk, len(seq)))
    n = sum(int(data[j][ix['Instructions Executed']]) for j in rng)
    th = sum(int(data[j][ix['Thread Instructions Executed']]) for j in rng)
    smp = sum(int(data[j][ix['# Samples']]) for j in rng)
    locs = [seq[j][1][1] for j in rng if seq[j][1] and seq[j][1][0].endswith('.cu')]
    print(f"SASS {k:5d}-{k + chunk:5d}  lines {min(locs) if locs else 0}-{max(locs) if locs else 0}  {100 * n / tot:5.1f}%  "
          f"{n / probes:5.2f} winstr/probe  {th / probes:6.1f} lane-instr/probe  avg threads {th / max(n, 1):4.1f}  samples {smp}")
