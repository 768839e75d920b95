#!/usr/bin/env python
"""Join an ncu source-page CSV (per-SASS-instruction counters) with nvdisasm -g line info of the
SAME build: dynamic warp instructions and stall samples per CUDA source line.
  ncu -i rep --page source --csv > src.csv ; cuobjdump -xelf all lib.so ; nvdisasm -g -c x.cubin > dis.txt
  sass_lines.py src.csv dis.txt <mangled kernel name prefix> <probes>"""
import collections, csv, re, sys
src_csv, dis_txt, kname, probes = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
lines = open(dis_txt).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("//--------------------- .text." + kname))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("//--------------------- ")), len(lines))
cur, seq = None, []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
    if m:
        seq.append((m.group(2).strip(), cur))
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
assert len(data) >= len(seq), (len(data), len(seq))
agg = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for (txt, loc), r in zip(seq, data):
    n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    agg[loc][0] += n; agg[loc][1] += s
    tot_i += n; tot_s += s
print(f"{len(seq)} SASS instructions, {tot_i/probes:.2f} warp instr / probe, {tot_s} samples")
src_cache = {}
for loc, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    text = ""
    if loc:
        f = "kmers.anno_b200/csrc/" + loc[0]
        try:
            src_cache.setdefault(f, open(f).read().split("\n"))
            text = src_cache[f][loc[1] - 1].strip()[:70]
        except OSError:
            pass
    print(f"{str(loc):28s} instr {100*n/tot_i:5.1f}% ({n/probes:5.2f}/probe) samples {100*s/max(tot_s,1):5.1f}%  {text}")
