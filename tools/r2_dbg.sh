#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_ops_attention_gpu.py -x -q -m gpu -s -k "test_global_attention and 0.0" > gpurun_out/g3_debug.log 2>&1
python - <<'PY'
import re, collections
seen = collections.OrderedDict()
for l in open("gpurun_out/g3_debug.log"):
    m = re.search(r"block=\((\d+),(\d+)\) thread=(\d+) bar=(\d+) parity=(\d+)", l)
    if not m: continue
    bx, by, t, bar, par = map(int, m.groups())
    key = (bx, by, t // 32, bar, par)
    seen[key] = seen.get(key, 0) + 1
blocks = sorted({(k[0], k[1]) for k in seen})
print("blocks with timeouts:", len(blocks), blocks[:12])
for k, c in list(seen.items())[:400]:
    if (k[0], k[1]) in blocks[:2] or k[0] == 10:
        print("block", k[:2], "warp", k[2], "bar idx", (k[3] - 1024 - 221184) // 8, "parity", k[4], "lanes", c)
PY
