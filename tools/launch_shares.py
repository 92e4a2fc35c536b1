"""Per-kernel shares of ONE step from an ncu launch list (--metrics gpu__time_duration.sum --csv): the launches between two
consecutive im2col16 launches (= one encoder + decoder pass).  usage: python tools/launch_shares.py <csv> [label]"""
import collections, csv, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = [(r[ki], float(r[vi].replace(",", "")) / 1e3) for r in rows[1:]]
idx = [i for i, (k, _) in enumerate(data) if "im2col" in k]
if len(idx) < 2:
    sys.exit("need two im2col16 launches in the list to delimit a step")
step = data[idx[-2]:idx[-1]]
agg = collections.OrderedDict()
for k, t in step:
    k = k.replace("void ", "").replace("<unnamed>::", "").replace("ullsam::", "")
    for cut in ("(CUtensorMap", "(const", "(Dec", "(float", "(int"):
        k = k.split(cut)[0]
    a = agg.setdefault(k[:64], [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(t for _, t in step)
print(f"{sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}: one step, {len(step)} launches, {tot / 1e3:.2f} ms serialised under ncu "
      "(cold caches, burst clocks: compare shares, not absolutes)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {t / 1e3:8.3f} ms {100 * t / tot:5.1f} %  {n:3d} x {t / n:8.1f} us  {k}")
