#!/usr/bin/env python
"""Stall totals and the most-sampled SASS instructions of an ncu report captured with --import-source on:
python tools/ncu_hot_sass.py <rep> [top=30]"""
import csv, subprocess, sys
rep = sys.argv[1]; top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))[1:]
h, body = rows[0], rows[1:]
s0, src, ex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
cols = {k: h.index(k) for k in h if k.startswith("stall_") and "Not Issued" not in k}
print("total samples", sum(int(r[s0] or 0) for r in body))
print("  ".join(f"{k[6:]}={sum(int(r[i] or 0) for r in body)}" for k, i in cols.items()))
order = {id(r): i for i, r in enumerate(body)}
for r in sorted(body, key=lambda r: -int(r[s0] or 0))[:top_n]:
    print(f"{order[id(r)]:5d} {r[s0]:>6s} ex {r[ex]:>7s} ", " ".join(f"{k[6:]}={r[i]}" for k, i in cols.items() if int(r[i] or 0) > 0), "|", r[src][:90])
