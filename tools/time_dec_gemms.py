"""The decoder's image-side GEMMs alone (CUDA events): cat projection (bf16 + pos), final cat, LN_RES."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ullsam_b200 import _lib, ops
M = 32 * 4096
g = torch.Generator(device="cuda").manual_seed(0)
def t(name, fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / reps * 1e3:8.1f} us")
keys_bf = torch.randn(M, 256, device="cuda", generator=g).bfloat16()
for N in (384, 256):
    w = (torch.randn(N, 256, device="cuda", generator=g) * 0.05).bfloat16(); b = torch.randn(N, device="cuda", generator=g)
    pos = torch.randn(4096, N, device="cuda", generator=g)
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    t(f"cat GEMM N={N} (bf16 + pos)", lambda: ops.gemm(keys_bf, w, b, epi=_lib.EPI_BF16, pos=pos, out=out))
    t(f"cat GEMM N={N} (bf16, no pos)", lambda: ops.gemm(keys_bf, w, b, epi=_lib.EPI_BF16, out=out))
ao = torch.randn(M, 128, device="cuda", generator=g).bfloat16()
w = (torch.randn(256, 128, device="cuda", generator=g) * 0.05).bfloat16(); b = torch.randn(256, device="cuda", generator=g)
keys = torch.randn(M, 256, device="cuda", generator=g); out2 = torch.empty(M, 256, dtype=torch.bfloat16, device="cuda")
lw = torch.ones(256, device="cuda"); lb = torch.zeros(256, device="cuda")
t("LN_RES (K=128, in place)", lambda: ops.gemm(ao, w, b, epi=_lib.EPI_LN_RES, out=keys, residual=keys, ln_w=lw, ln_b=lb, ln_eps=1e-5, out2=out2))
