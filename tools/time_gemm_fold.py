"""Encoder GEMMs with the LayerNorm fold vs the plain epilogues + the LayerNorm kernel they replace (CUDA events, B=32 ViT-B
shapes, working set > L2).  usage: python tools/time_gemm_fold.py [vit_l]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ullsam_b200 import _lib, ops

def bench(name, fn, reps=8):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / reps * 1e3:8.1f} us")

vitl = len(sys.argv) > 1 and sys.argv[1] == "vit_l"
D, MD, heads = (1024, 4096, 16) if vitl else (768, 3072, 12)
M = 131072
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, D, device="cuda", generator=g)
xb = x.bfloat16()
st = torch.zeros(M, _lib.STAT_SLOTS, 2, device="cuda")
xs = x.reshape(M, D // 128, 128)
st[:, : D // 128, 0] = xs.sum(-1); st[:, : D // 128, 1] = (xs * xs).sum(-1)
del xs
gam = torch.ones(D, device="cuda"); bet = torch.zeros(D, device="cuda")
bench("layernorm kernel", lambda: ops.layernorm(x.reshape(32, 64, 64, D), gam, bet, 1e-6))
def mk(N, K):
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.03).bfloat16()
    return w, torch.randn(N, device="cuda", generator=g), w.float().sum(1).contiguous()
w, b, s = mk(3 * D, D)
q = torch.empty(32, heads, 4096, 64, dtype=torch.bfloat16, device="cuda"); k = torch.empty_like(q); v = torch.empty_like(q)
bench("qkv", lambda: ops.gemm(xb, w, b, epi=_lib.EPI_QKV, qkv=(q, k, v), T=4096, heads=heads))
bench("qkv fold", lambda: ops.gemm(xb, w, b, epi=_lib.EPI_QKV_FOLD, qkv=(q, k, v), T=4096, heads=heads, fold_s=s, row_stats=st))
del q, k, v
w, b, s = mk(MD, D)
hid = torch.empty(M, MD, dtype=torch.bfloat16, device="cuda")
bench("lin1 gelu", lambda: ops.gemm(xb, w, b, epi=_lib.EPI_BF16, act=_lib.ACT_GELU, out=hid))
bench("lin1 gelu fold", lambda: ops.gemm(xb, w, b, epi=_lib.EPI_GELU_FOLD, out=hid, fold_s=s, row_stats=st))
w, b, s = mk(D, D)
bench("proj res", lambda: ops.gemm(xb, w, b, epi=_lib.EPI_RES_F32, out=x, residual=x))
xo = torch.empty_like(xb)
bench("proj res + stats", lambda: ops.gemm(xb, w, b, epi=_lib.EPI_RES_STATS, out=x, residual=x, out2=xo, row_stats=st))
w, b, s = mk(D, MD)
bench("lin2 res", lambda: ops.gemm(hid, w, b, epi=_lib.EPI_RES_F32, out=x, residual=x))
bench("lin2 res + stats", lambda: ops.gemm(hid, w, b, epi=_lib.EPI_RES_STATS, out=x, residual=x, out2=xo, row_stats=st))
