#!/usr/bin/env python
"""Top stall sites of one kernel from `ncu --page source --csv --print-source sass` (first kernel in the file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
iS = hdr.index("Warp Stall Sampling (All Samples)")
iSrc = hdr.index("Source")
iEx = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS] or 0) for r in body)
print(f"total samples {tot}, instructions {len(body)}")
agg = {}
for i in stall_cols:
    agg[hdr[i]] = sum(int(r[i] or 0) for r in body)
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(body)), key=lambda k: -int(body[k][iS] or 0))[:top]
for k in sorted(order):
    r = body[k]
    why = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {int(r[iS] or 0):6d} {100.0 * int(r[iS] or 0) / max(tot, 1):5.1f}%  ex={r[iEx]:>8s}  {r[iSrc][:70]:70s} {why}")
