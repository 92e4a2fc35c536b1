"""Time the tcgen05 GEMM op alone for the encoder's shapes and epilogues (CUDA events, working set > L2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ullsam_b200 import _lib, ops

def bench(name, fn, flops, reps=8):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:44s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.0f} TF/s")

M = 131072
g = torch.Generator(device="cuda").manual_seed(0)
def mk(N, K):
    a = (torch.randn(M, K, device="cuda", generator=g)).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.03).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    return a, w, b
for N, K in ((2304, 768), (3072, 768), (768, 768), (768, 3072)):
    a, w, b = mk(N, K)
    fl = 2.0 * M * N * K
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    bench(f"N={N} K={K} bf16", lambda: ops.gemm(a, w, b, epi=_lib.EPI_BF16, out=out), fl)
    bench(f"N={N} K={K} bf16 + GELU", lambda: ops.gemm(a, w, b, epi=_lib.EPI_BF16, act=_lib.ACT_GELU, out=out), fl)
    if N == 2304:
        q = torch.empty(32, 12, 4096, 64, dtype=torch.bfloat16, device="cuda"); k = torch.empty_like(q); v = torch.empty_like(q)
        bench(f"N={N} K={K} qkv split", lambda: ops.gemm(a, w, b, epi=_lib.EPI_QKV, qkv=(q, k, v), T=4096, heads=12), fl)
    if N == 768:
        x = torch.randn(M, N, device="cuda", generator=g)
        bench(f"N={N} K={K} +residual fp32", lambda: ops.gemm(a, w, b, epi=_lib.EPI_RES_F32, out=x, residual=x), fl)
    del a, w, b, out
