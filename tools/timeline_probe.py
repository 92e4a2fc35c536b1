import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ullsam_b200 import ops, _lib
grid, G, heads = 64, 32, 12
T = grid * grid
q = torch.randn(G, heads, T, 64, device="cuda").bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q)
rh = torch.randn(127, 64, device="cuda") * 0.1; rw = torch.randn(127, 64, device="cuda") * 0.1
for _ in range(3): ops.attention(q, k, v, rh, rw, grid)
torch.cuda.synchronize()
buf = (C.c_uint * (4 * 16 * 8))()
lib = _lib.load()
lib.ullsam_probe_timeline.argtypes = [C.c_void_p]
print("rc", lib.ullsam_probe_timeline(buf))
import numpy as np
a = np.array(buf, dtype=np.int64).reshape(4, 16, 8)
t0 = a[0, 0, 0]
names = {0: "softmax qt0 [wait_s, got_s, ldtm, xch, passB, st, o_done]", 1: "softmax qt1", 2: "mma item qt0 [wait_p, got_p, pv, qk]", 3: "mma item qt1"}
for role in range(4):
    print(names[role])
    for j in range(16):
        print("  tile", j + 8, [(int(x - t0) if x else None) for x in a[role, j, :7]])
