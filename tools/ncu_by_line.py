#!/usr/bin/env python
"""Executed warp instructions per source line: joins `ncu --page source --csv --print-source sass` (per-instruction
counts) with `nvdisasm -g` of the same cubin (file/line per instruction offset).

    cuobjdump -xelf all build/kernels_tuned.o; nvdisasm -g kernels_tuned.sm_100a.cubin > kt.sass
    python tools/ncu_by_line.py gpurun_out/x_sass.csv kt.sass '<mangled kernel name substring>' [top]
"""
import csv, re, sys, collections
sass_csv, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
iEx, iSrc, iSm = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
base = int(body[0][0], 16)
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
cur = ("?", 0)
loc = {}
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m:
        loc[int(m.group(1), 16)] = (cur, m.group(2))
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot = 0
FP64 = ("DFMA", "DMUL", "DADD", "DSETP")
for r in body:
    off = int(r[0], 16) - base
    ex = int(float(r[iEx] or 0))
    (fl, op) = loc.get(off, (("?", 0), "?"))
    a = agg[fl]
    a[0] += ex
    opn = next((w for w in r[iSrc].split() if not w.startswith("@")), "?").split(".")[0]
    if opn in FP64:
        a[1] += ex
    a[2] += int(r[iSm] or 0)
    a[3][opn] += ex
    tot += ex
print(f"total warp instructions {tot:,}")
print(f"{'file:line':34s} {'warp insts':>14s} {'%':>6s} {'fp64':>14s} {'samples':>8s}  top ops")
for fl, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    ops = ", ".join(f"{k} {v / max(a[0], 1):.0%}" for k, v in a[3].most_common(3))
    print(f"{fl[0] + ':' + str(fl[1]):34s} {a[0]:14,d} {100 * a[0] / tot:6.2f} {a[1]:14,d} {a[2]:8d}  {ops}")
