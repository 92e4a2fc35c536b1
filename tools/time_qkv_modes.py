"""QKV GEMM alone in its three row layouts (CUDA events): global (T=4096), windowed on pre-partitioned rows (4900 per image,
the round-1 path), windowed with window_partition on the store (4096 rows per image + pad rows from the epilogue)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ullsam_b200 import _lib, ops
B, heads, D = 32, 12, 768
g = torch.Generator(device="cuda").manual_seed(0)
w = (torch.randn(3 * D, D, device="cuda", generator=g) * 0.03).bfloat16()
b = torch.randn(3 * D, device="cuda", generator=g)
a_tok = torch.randn(B * 4096, D, device="cuda", generator=g).bfloat16()
a_win = torch.randn(B * 4900, D, device="cuda", generator=g).bfloat16()
q = torch.empty(B * 25, heads, 196, 64, dtype=torch.bfloat16, device="cuda"); k = torch.empty_like(q); v = torch.empty_like(q)
qg, kg, vg = (t.view(-1)[:B * heads * 4096 * 64].view(B, heads, 4096, 64) for t in (q, k, v))
modes = {
    "global   (M=131072, T=4096)": lambda: ops.gemm(a_tok, w, b, epi=_lib.EPI_QKV, qkv=(qg, kg, vg), T=4096, heads=heads),
    "windowed (M=156800, rows pre-partitioned)": lambda: ops.gemm(a_win, w, b, epi=_lib.EPI_QKV, qkv=(q, k, v), T=196, heads=heads),
    "windowed (M=131072, partition on store)": lambda: ops.gemm(a_tok, w, b, epi=_lib.EPI_QKV, qkv=(q, k, v), T=196, heads=heads, ln_partition=True),
}
for rnd in range(2):
    for name, fn in modes.items():
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 20 * 1000:.1f} us")
