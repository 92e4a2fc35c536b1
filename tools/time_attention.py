"""Time the fused attention op alone (CUDA events, L2-exceeding working set)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ullsam_b200 import ops
for grid, G, heads in ((64, 32, 12), (14, 800, 12)):
    T = grid * grid
    q = torch.randn(G, heads, T, 64, device="cuda").bfloat16()
    k = torch.randn(G, heads, T, 64, device="cuda").bfloat16()
    v = torch.randn(G, heads, T, 64, device="cuda").bfloat16()
    rh = torch.randn(2 * grid - 1, 64, device="cuda") * 0.1
    rw = torch.randn(2 * grid - 1, 64, device="cuda") * 0.1
    for _ in range(3): ops.attention(q, k, v, rh, rw, grid)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.attention(q, k, v, rh, rw, grid)
    e1.record(); torch.cuda.synchronize()
    print(f"grid {grid}: {e0.elapsed_time(e1) / 10:.3f} ms per launch (B=32 equivalent)")
