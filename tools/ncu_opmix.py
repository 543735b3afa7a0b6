#!/usr/bin/env python
"""Opcode mix + LSU wavefronts of one kernel from `ncu --page source --csv --print-source sass`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
col = {h: i for i, h in enumerate(hdr)}
def num(r, name):
    try:
        return float(r[col[name]] or 0)
    except (KeyError, ValueError):
        return 0.0
ops = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
for r in body:
    src = r[col["Source"]].split()
    op = next((w for w in src if not w.startswith("@")), "?").split(".")[0]
    full = next((w for w in src if not w.startswith("@")), "?")
    key = full if op in ("LDS", "STS", "LDG", "STG", "LDGSTS", "RED", "REDG", "ATOMG") else op
    o = ops[key]
    o[0] += num(r, "Instructions Executed"); o[1] += num(r, "L1 Wavefronts Shared"); o[2] += num(r, "L1 Wavefronts Shared Ideal")
    o[3] += num(r, "L1 Tag Requests Global"); o[4] += num(r, "L2 Theoretical Sectors Global"); o[5] += num(r, "Warp Stall Sampling (All Samples)")
tot = sum(o[0] for o in ops.values()); ts = sum(o[5] for o in ops.values())
print(f"total warp instructions {tot:,.0f}  samples {ts:,.0f}")
print(f"{'op':24s} {'warp insts':>14s} {'%':>6s} {'smem wf':>12s} {'ideal':>12s} {'L1 tag req':>12s} {'L2 sectors':>12s} {'samples%':>8s}")
for k, o in sorted(ops.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 28]:
    print(f"{k:24s} {o[0]:14,.0f} {100*o[0]/tot:6.1f} {o[1]:12,.0f} {o[2]:12,.0f} {o[3]:12,.0f} {o[4]:12,.0f} {100*o[5]/max(ts,1):8.1f}")
