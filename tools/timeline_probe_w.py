import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ullsam_b200 import ops, _lib
grid, G, heads = 14, 800, 12
T = grid * grid
q = torch.randn(G, heads, T, 64, device="cuda").bfloat16(); k = torch.randn_like(q); v = torch.randn_like(q)
rh = torch.randn(27, 64, device="cuda") * 0.1; rw = torch.randn(27, 64, device="cuda") * 0.1
for _ in range(3): ops.attention(q, k, v, rh, rw, grid)
torch.cuda.synchronize()
buf = (C.c_uint * 64)()
lib = _lib.load()
lib.ullsam_probe_timeline.argtypes = [C.c_void_p]
print("rc", lib.ullsam_probe_timeline(buf))
t = list(buf)
names = ["start", "after alloc+sync", "pro_mma_done", "gather done", "S0 ready", "P0 arrived", None, None, "S1 ready", "P1 arrived", None, None, "o_done", "epilogue stored", "after final sync", "end"]
for i, n in enumerate(names):
    if n: print(f"{n:20s} {t[i] - t[0]:8d}")
