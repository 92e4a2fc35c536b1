"""ViT-L QKV GEMM alone (M = 32 * 4096, N = 3072, K = 1024), a few launches, for ncu and timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ullsam_b200 import _lib, ops
M, N, K, heads = 131072, 3072, 1024, 16
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
w = (torch.randn(N, K, device="cuda", generator=g) * 0.03).bfloat16()
b = torch.randn(N, device="cuda", generator=g)
q = torch.empty(32, heads, 4096, 64, dtype=torch.bfloat16, device="cuda"); k = torch.empty_like(q); v = torch.empty_like(q)
fn = lambda: ops.gemm(a, w, b, epi=_lib.EPI_QKV, qkv=(q, k, v), T=4096, heads=heads)
for _ in range(4): fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8): fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 8
print(f"ViT-L qkv GEMM {M}x{N}x{K}: {ms * 1e3:.1f} us, {2.0 * M * N * K / ms / 1e9:.0f} TF/s")
