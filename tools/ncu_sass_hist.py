#!/usr/bin/env python
"""Histogram of executed warp instructions by opcode (and top stall PCs) from `ncu --page source --csv`."""
import csv, sys, collections
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci, si, ei = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
op = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
lines = []
for k, r in enumerate(rows[hi + 1:]):
    if len(r) <= ei: continue
    ins = r[ci].strip()
    toks = ins.split()
    name = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    name = name.split(".")[0]
    n = int(r[ei] or 0); s = int(r[si] or 0)
    op[name] += n; samp[name] += s; tot += n; tots += s
    lines.append((k, n, s, ins))
print(f"total warp instructions {tot:,}  samples {tots:,}")
for name, n in op.most_common(22):
    print(f"  {name:10s} {n:>14,} {100*n/tot:5.1f}%   samples {100*samp[name]/max(tots,1):5.1f}%")
if len(sys.argv) > 2:
    # executed-instruction profile along the program: buckets of 64 SASS lines
    B = int(sys.argv[2])
    for b in range(0, len(lines), B):
        chunk = lines[b:b + B]
        n = sum(c[1] for c in chunk); s = sum(c[2] for c in chunk)
        print(f"  lines {b:5d}-{b+B:5d}  insts {100*n/tot:5.1f}%  samples {100*s/max(tots,1):5.1f}%   {chunk[0][3][:60]}")
