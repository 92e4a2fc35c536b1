"""Run each fused kernel several times on identical inputs and report run-to-run differences (race detector)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ullsam_b200 import ops

def inputs(G, heads, grid, seed, rel_scale=0.25):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    T = grid * grid
    q = (torch.randn(G, heads, T, 64, device="cuda", generator=gen) * 1.2).bfloat16()
    k = (torch.randn(G, heads, T, 64, device="cuda", generator=gen) * 1.2).bfloat16()
    v = torch.randn(G, heads, T, 64, device="cuda", generator=gen).bfloat16()
    rh = (torch.randn(2 * grid - 1, 64, device="cuda", generator=gen) * rel_scale).bfloat16().float()
    rw = (torch.randn(2 * grid - 1, 64, device="cuda", generator=gen) * rel_scale).bfloat16().float()
    return q, k, v, rh, rw

for grid, G, heads in ((64, 8, 12), (14, 200, 12)):
    q, k, v, rh, rw = inputs(G, heads, grid, 5)
    ref = ops.attention(q, k, v, rh, rw, grid)
    torch.cuda.synchronize()
    for it in range(6):
        out = ops.attention(q, k, v, rh, rw, grid)
        torch.cuda.synchronize()
        d = (out.float() - ref.float()).abs()
        nbad = int((d > 0).sum())
        print(f"grid {grid} run {it}: differing elements {nbad} max diff {d.max().item():.3e}", flush=True)
        if nbad:
            idx = (d > 0).nonzero()
            print("   first diffs (b,y,x,c):", idx[:6].tolist(), " rows affected:", len(set((int(a), int(b), int(c)) for a, b, c, _ in idx[:2000].tolist())))
