"""One encoder GEMM alone, a few launches, for ncu: python tools/time_gemm_one.py proj|lin2|qkv|lin1"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ullsam_b200 import _lib, ops
which = sys.argv[1] if len(sys.argv) > 1 else "proj"
M = 131072
N, K = {"proj": (768, 768), "lin2": (768, 3072), "qkv": (2304, 768), "lin1": (3072, 768)}[which]
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
w = (torch.randn(N, K, device="cuda", generator=g) * 0.03).bfloat16()
b = torch.randn(N, device="cuda", generator=g)
if which in ("proj", "lin2"):
    x = torch.randn(M, N, device="cuda", generator=g)
    fn = lambda: ops.gemm(a, w, b, epi=_lib.EPI_RES_F32, out=x, residual=x)
elif which == "qkv":
    q = torch.empty(32, 12, 4096, 64, dtype=torch.bfloat16, device="cuda"); k = torch.empty_like(q); v = torch.empty_like(q)
    fn = lambda: ops.gemm(a, w, b, epi=_lib.EPI_QKV, qkv=(q, k, v), T=4096, heads=12)
else:
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    fn = lambda: ops.gemm(a, w, b, epi=_lib.EPI_BF16, act=_lib.ACT_GELU, out=out)
for _ in range(4):
    fn()
torch.cuda.synchronize()
