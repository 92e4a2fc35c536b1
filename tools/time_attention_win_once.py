"""Run the windowed fused attention op twice (for timeline-instrumented debug builds that printf from the kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ullsam_b200 import ops
grid, G, heads = 14, 800, 12
T = grid * grid
q = torch.randn(G, heads, T, 64, device="cuda").bfloat16()
k = torch.randn(G, heads, T, 64, device="cuda").bfloat16()
v = torch.randn(G, heads, T, 64, device="cuda").bfloat16()
rh = torch.randn(2 * grid - 1, 64, device="cuda") * 0.1
rw = torch.randn(2 * grid - 1, 64, device="cuda") * 0.1
for _ in range(2):
    ops.attention(q, k, v, rh, rw, grid)
torch.cuda.synchronize()
