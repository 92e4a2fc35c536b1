#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples + per-reason totals.
usage: ncu -i rep --page source --csv | python tools/ncu_stalls.py [kernel-substring] [topN]"""
import csv
import sys

sub = sys.argv[1] if len(sys.argv) > 1 else ""
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(sys.stdin))
kernels, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kernels.append(cur)
    elif r and r[0] == "Address":
        hdr = r
    elif cur is not None and hdr is not None and len(r) == len(hdr):
        cur["rows"].append(r)
for k in kernels:
    if sub not in k["name"]:
        continue
    print("=" * 100)
    print(k["name"][:140])
    si = hdr.index("# Samples")
    ei = hdr.index("Instructions Executed")
    reasons = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si]) for r in k["rows"])
    print("total samples", tot, " instructions", len(k["rows"]))
    rt = {hdr[i]: sum(int(r[i]) for r in k["rows"]) for i in reasons}
    for name, v in sorted(rt.items(), key=lambda x: -x[1])[:10]:
        print(f"  {name:28s} {v:8d} {100.0 * v / max(tot, 1):5.1f}%")
    order = sorted(range(len(k["rows"])), key=lambda j: -int(k["rows"][j][si]))[:top]
    for j in sorted(order):
        r = k["rows"][j]
        why = sorted(((int(r[i]), hdr[i]) for i in reasons), reverse=True)[:2]
        print(f"  #{j:5d} {int(r[si]):7d} ({100.0 * int(r[si]) / max(tot, 1):4.1f}%) exec={r[ei]:>9s} {r[1].strip()[:70]:70s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
    break
