"""The attention-projection GEMM alone (M=131072, N=K=768, fp32 residual epilogue): a few launches for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ullsam_b200 import _lib, ops
M, N, K = 131072, 768, int(sys.argv[1]) if len(sys.argv) > 1 else 768
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
w = (torch.randn(N, K, device="cuda", generator=g) * 0.03).bfloat16()
b = torch.randn(N, device="cuda", generator=g)
x = torch.randn(M, N, device="cuda", generator=g)
for _ in range(4):
    ops.gemm(a, w, b, epi=_lib.EPI_RES_F32, out=x, residual=x)
torch.cuda.synchronize()
