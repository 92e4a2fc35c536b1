"""tcgen05 GEMM epilogues vs a plain torch fp32 computation of the same op (floating-point kernel:
tolerance = bf16 operand rounding only, accumulation is fp32 on both sides)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(M, N, K, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    return a, w, b


def _ref(a, w, b):
    return a.float() @ w.float().t() + b


def _close(got, ref, tol, what):
    err = (got.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    print(f"{what}: max|err|={err:.3e} scale={scale:.3e}")
    assert err <= tol * scale, f"{what}: max|err| {err} vs scale {scale}"


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 128, 128), (384, 768, 768), (1000, 2304, 768),
                                   (4900, 768, 3072), (131072 // 8, 3072, 768)])
def test_gemm_bias_bf16(M, N, K):
    from ullsam_b200 import _lib, ops

    a, w, b = _mk(M, N, K)
    out = ops.gemm(a, w, b, epi=_lib.EPI_BF16)
    torch.cuda.synchronize()
    _close(out, _ref(a, w, b), 8e-3, f"bias_bf16 {M}x{N}x{K}")


def test_gemm_gelu_relu():
    from ullsam_b200 import _lib, ops

    a, w, b = _mk(512, 3072, 768, 1)
    ref = _ref(a, w, b)
    out = ops.gemm(a, w, b, epi=_lib.EPI_BF16, act=_lib.ACT_GELU)
    _close(out, torch.nn.functional.gelu(ref), 8e-3, "gelu")
    out = ops.gemm(a, w, b, epi=_lib.EPI_F32, act=_lib.ACT_RELU)
    _close(out, torch.relu(ref), 1e-4, "relu_f32")


def test_gemm_residual_inplace():
    from ullsam_b200 import _lib, ops

    a, w, b = _mk(640, 768, 3072, 2)
    res = torch.randn(640, 768, device="cuda")
    ref = _ref(a, w, b) + res
    x = res.clone()
    ops.gemm(a, w, b, epi=_lib.EPI_RES_F32, out=x, residual=x)
    _close(x, ref, 1e-4, "residual")


@pytest.mark.parametrize("M,N,K", [(4096, 768, 768), (2 * 4096 + 300, 1024, 1024), (4096 + 17, 768, 3072)])
def test_gemm_residual_by_tma_reduce_add(M, N, K):
    """EPI_RES_RED: x += A W^T + b with the add done by TMA reduce-add stores (the residual stream never enters the SM).
    Same two roundings as EPI_RES_F32 (fp32(acc + b), then + x), so the two must agree bit for bit; rows beyond M of the
    last 32-row box are clipped by the tensor map."""
    from ullsam_b200 import _lib, ops

    a, w, b = _mk(M, N, K, 31)
    res = torch.randn(M + 64, N, device="cuda")          # 64 guard rows behind the matrix: must stay untouched
    x, y = res.clone(), res.clone()
    ops.gemm(a, w, b, epi=_lib.EPI_RES_RED, out=x[:M], residual=x[:M])
    ops.gemm(a, w, b, epi=_lib.EPI_RES_F32, out=y[:M], residual=y[:M])
    torch.cuda.synchronize()
    _close(x[:M], _ref(a, w, b) + res[:M], 1e-4, f"res_red {M}x{N}x{K}")
    assert torch.equal(x[M:], res[M:]), "rows behind the matrix were written"
    assert torch.equal(x[:M], y[:M]), "reduce-add and in-SM add must agree bit for bit"


@pytest.mark.parametrize("partition", [False, True])
@pytest.mark.parametrize("K", [768, 3072])
def test_gemm_residual_with_fused_layernorm(K, partition):
    """EPI_RES_LN: x += A W^T + b in place AND bf16 LayerNorm(x) of the updated rows (Block.norm2 / next norm1,
    image_encoder.py:166-182), in token order or in window_partition order (image_encoder.py:243-264)."""
    from ullsam_b200 import _lib, ops

    B, dim = 3, 768
    M = B * 4096   # includes a tail: 12288 rows = 48 row blocks of 256, dealt M-major to the CTA pairs
    a, w, b = _mk(M, dim, K, 21)
    res = torch.randn(M, dim, device="cuda") * 3.0 + 0.5   # non-zero mean, larger scale than the GEMM term
    gamma = 1.0 + 0.1 * torch.randn(dim, device="cuda")
    beta = 0.1 * torch.randn(dim, device="cuda")
    x_ref = _ref(a, w, b) + res
    ln_ref = torch.nn.functional.layer_norm(x_ref, (dim,), gamma, beta, 1e-6)
    x = res.clone()
    rows = B * 25 * 196 if partition else M
    xn = torch.full((rows, dim), 7.0, dtype=torch.bfloat16, device="cuda")   # sentinel: pad rows must stay untouched
    ops.gemm(a, w, b, epi=_lib.EPI_RES_LN, out=x, residual=x, ln_w=gamma, ln_b=beta, ln_eps=1e-6, out2=xn,
             ln_partition=partition)
    _close(x, x_ref, 1e-4, f"res+ln K={K}: residual stream")
    if partition:
        # reference window partition (pad 64 -> 70, 5x5 windows of 14x14), zero pad rows replaced by the sentinel
        g = torch.full((B, 70, 70, dim), 7.0, device="cuda")
        g[:, :64, :64] = ln_ref.view(B, 64, 64, dim)
        ln_ref = g.view(B, 5, 14, 5, 14, dim).permute(0, 1, 3, 2, 4, 5).reshape(rows, dim)
    _close(xn, ln_ref, 1.2e-2, f"res+ln K={K} partition={partition}: normalised rows")


def test_gemm_patch_pos():
    from ullsam_b200 import _lib, ops

    a, w, b = _mk(8192, 768, 768, 3)
    pos = torch.randn(4096, 768, device="cuda") * 0.02
    out = ops.gemm(a, w, b, epi=_lib.EPI_PATCH, pos=pos)
    ref = _ref(a, w, b) + pos.repeat(2, 1)
    _close(out, ref, 1e-4, "patch+pos")


@pytest.mark.parametrize("T,G", [(196, 50), (4096, 2)])
def test_gemm_qkv_split(T, G):
    from ullsam_b200 import _lib, ops

    heads, dim = 12, 768
    a, w, b = _mk(G * T, 3 * dim, dim, 4)
    q = torch.zeros(G, heads, T, 64, dtype=torch.bfloat16, device="cuda")
    k, v = torch.zeros_like(q), torch.zeros_like(q)
    ops.gemm(a, w, b, epi=_lib.EPI_QKV, qkv=(q, k, v), T=T, heads=heads)
    ref = _ref(a, w, b).reshape(G, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
    for name, got, r in (("q", q, ref[0]), ("k", k, ref[1]), ("v", v, ref[2])):
        _close(got, r, 8e-3, f"qkv[{name}] T={T}")


def test_gemm_qkv_key_norm_maximum():
    """EPI_QKV with kmax_sq: the epilogue also folds max_t |k[g][head][t]|^2 into a [G][heads] buffer (the a-priori score
    bound of the global attention, attention.cu) — the values are the fp32 ones before the bf16 rounding of the store."""
    from ullsam_b200 import _lib, ops

    T, G, heads, dim = 4096, 2, 12, 768
    a, w, b = _mk(G * T - 100, 3 * dim, dim, 6)          # the last group is short: its missing rows must not contribute
    q = torch.zeros(G, heads, T, 64, dtype=torch.bfloat16, device="cuda")
    k, v = torch.zeros_like(q), torch.zeros_like(q)
    kmax = torch.zeros(G, heads, device="cuda")
    ops.gemm(a, w, b, epi=_lib.EPI_QKV, qkv=(q, k, v), T=T, heads=heads, kmax_sq=kmax)
    torch.cuda.synchronize()
    kk = (_ref(a, w, b)[:, dim:2 * dim]).reshape(-1, heads, 64)
    rows = torch.arange(kk.shape[0], device="cuda") // T
    want = torch.stack([(kk[rows == g] ** 2).sum(-1).amax(0) for g in range(G)])
    _close(kmax, want, 1e-4, "max |k|^2 per (group, head)")
    assert (kmax * 1.004 ** 2 >= (k.float() ** 2).sum(-1).amax(-1)).all(), "must bound the norms of the stored bf16 rows"


def test_gemm_qkv_window_partition_on_store():
    """The windowed blocks' QKV projection (image_encoder.py:168-172, 227-229): the reference zero-pads the normalised
    64x64 map to 70x70, partitions it into 25 windows of 14x14 and projects all 4900 tokens per image.  Here the GEMM runs
    on the 4096 image tokens; its epilogue scatters the rows into the windows and writes k = v = bias for the 804 padding
    tokens (q rows of padding tokens are never used: window_unpartition crops them)."""
    from ullsam_b200 import _lib, ops

    B, heads, dim = 2, 12, 768
    a, w, b = _mk(B * 4096, 3 * dim, dim, 5)
    q = torch.full((B * 25, heads, 196, 64), 7.0, dtype=torch.bfloat16, device="cuda")
    k, v = q.clone(), q.clone()
    ops.gemm(a, w, b, epi=_lib.EPI_QKV, qkv=(q, k, v), T=196, heads=heads, ln_partition=True)
    # reference: pad the INPUT with zero rows, partition, project (bias included)
    x = a.float().reshape(B, 64, 64, dim)
    xp = torch.nn.functional.pad(x, (0, 0, 0, 6, 0, 6))
    win = xp.reshape(B, 5, 14, 5, 14, dim).permute(0, 1, 3, 2, 4, 5).reshape(B * 25 * 196, dim)
    ref = (win @ w.float().t() + b).reshape(B * 25, 196, 3, heads, 64).permute(2, 0, 3, 1, 4)
    valid = torch.nn.functional.pad(torch.ones(B, 64, 64, device="cuda"), (0, 6, 0, 6))
    valid = valid.reshape(B, 5, 14, 5, 14).permute(0, 1, 3, 2, 4).reshape(B * 25, 1, 196, 1).bool()
    for name, got, r in (("k", k, ref[1]), ("v", v, ref[2])):
        _close(got, r, 8e-3, f"windowed qkv[{name}]")
    _close(torch.where(valid, q.float(), ref[0]), ref[0], 8e-3, "windowed qkv[q] (real tokens)")
    assert (q.float()[(~valid).expand_as(q)] == 7.0).all(), "q rows of padding tokens must stay untouched"


# ---- LayerNorm folded into the GEMMs on either side of it (the encoder's default path; Block.norm1 / norm2,
# image_encoder.py:166-182):  LN(x) W^T + b = rstd * (x (gamma*W)^T - mean * s) + c ----

def _row_stats(x):
    """[M, D] fp32 -> the producer epilogues' statistics buffer [M, 8, 2]: (sum, sum of squares) per 128-column slice."""
    from ullsam_b200 import _lib
    M, D = x.shape
    st = torch.zeros(M, _lib.STAT_SLOTS, 2, device=x.device)
    xs = x.reshape(M, D // 128, 128)
    st[:, : D // 128, 0] = xs.sum(-1)
    st[:, : D // 128, 1] = (xs * xs).sum(-1)
    return st


@pytest.mark.parametrize("M,N,K,patch", [(4096, 768, 768, False), (2 * 4096 + 300, 1024, 4096, False), (300, 768, 768, False),
                                         (4096, 768, 768, True)])
def test_gemm_residual_with_row_statistics(M, N, K, patch):
    """EPI_RES_STATS / EPI_PATCH_STATS: the fp32 result is the plain epilogue's, out2 is its bf16 copy and row_stats holds
    the partial sums of every 128-column slice (computed from the fp32 values)."""
    from ullsam_b200 import _lib, ops

    a, w, b = _mk(M, N, K, 11)
    st = torch.full((M, _lib.STAT_SLOTS, 2), -7.0, device="cuda")
    xb = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    if patch:
        pos = torch.randn(4096, N, device="cuda")
        ref = _ref(a, w, b) + pos.repeat(M // 4096, 1)
        x = ops.gemm(a, w, b, epi=_lib.EPI_PATCH_STATS, pos=pos, out2=xb, row_stats=st)
        plain = ops.gemm(a, w, b, epi=_lib.EPI_PATCH, pos=pos)
    else:
        res = torch.randn(M, N, device="cuda") * 2 + 0.3
        ref = _ref(a, w, b) + res
        x = res.clone()
        ops.gemm(a, w, b, epi=_lib.EPI_RES_STATS, out=x, residual=x, out2=xb, row_stats=st)
        plain = res.clone()
        ops.gemm(a, w, b, epi=_lib.EPI_RES_F32, out=plain, residual=plain)
    torch.cuda.synchronize()
    _close(x, ref, 1e-4, "res_stats out")
    assert torch.equal(x, plain), "the statistics epilogue must not change the fp32 result"
    assert torch.equal(xb, x.bfloat16()), "out2 is the bf16 rounding of out"
    want = _row_stats(x)
    ns = N // 128
    _close(st[:, :ns, 0], want[:, :ns, 0], 1e-5, "row sums")
    _close(st[:, :ns, 1], want[:, :ns, 1], 1e-5, "row sums of squares")
    assert (st[:, ns:] == -7.0).all(), "slots beyond N / 128 are not written"


def _fold_case(M, N, D, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    # a residual-stream-like input: per-channel scales, a few large channels, a non-zero row mean
    x = torch.randn(M, D, device="cuda", generator=g) * (0.5 + torch.rand(D, device="cuda", generator=g)) + 0.2
    x[:, 5] *= 20.0
    x[:, D - 3] += 8.0
    gamma = 1.0 + 0.3 * torch.randn(D, device="cuda", generator=g)
    beta = 0.2 * torch.randn(D, device="cuda", generator=g)
    w = torch.randn(N, D, device="cuda", generator=g) * 0.05
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    ref = torch.nn.functional.layer_norm(x, (D,), gamma, beta, 1e-6) @ w.t() + b
    wf = (gamma * w).bfloat16()
    _fold_case.affine = (gamma, beta)
    return x, w, b, ref, wf, wf.float().sum(1).contiguous(), (w @ beta + b).contiguous()


@pytest.mark.parametrize("M,D", [(4096, 768), (4096 + 200, 1024), (200, 768)])
def test_gemm_gelu_with_folded_layernorm(M, D):
    """EPI_GELU_FOLD = GELU(LayerNorm(x) W^T + b) computed from the bf16 copy of x, gamma*W and the row statistics;
    compared with fp32 torch and with the unfolded CUDA path (LayerNorm kernel -> GEMM) it replaces."""
    from ullsam_b200 import _lib, ops

    N = 3072
    x, w, b, ref, wf, s, c = _fold_case(M, N, D, 21)
    out = ops.gemm(x.bfloat16(), wf, c, epi=_lib.EPI_GELU_FOLD, fold_s=s, row_stats=_row_stats(x), ln_eps=1e-6)
    torch.cuda.synchronize()
    gref = torch.nn.functional.gelu(ref)
    _close(out, gref, 1.2e-2, f"gelu_fold M={M} D={D}")
    e_fold = ((out.float() - gref).norm() / gref.norm()).item()
    print(f"gelu_fold rel-L2 vs fp32: {e_fold:.3e}")
    assert e_fold <= 6e-3
    if M % 4096 == 0:  # the path it replaces: LayerNorm kernel (fp32 -> bf16), then the plain GELU GEMM
        gamma, beta = _fold_case.affine
        xn = ops.layernorm(x.reshape(M // 4096, 64, 64, D).contiguous(), gamma, beta, 1e-6).reshape(M, D)
        plain = ops.gemm(xn, w.bfloat16(), b, epi=_lib.EPI_BF16, act=_lib.ACT_GELU)
        e_plain = ((plain.float() - gref).norm() / gref.norm()).item()
        print(f"          unfolded path vs fp32: {e_plain:.3e}")
        assert e_fold <= 1.5 * e_plain + 1e-3


@pytest.mark.parametrize("partition", [False, True])
def test_gemm_qkv_with_folded_layernorm(partition):
    """EPI_QKV_FOLD: head split (and window_partition on the store) of LayerNorm(x) W^T + b; the k / v rows of the
    window-padding tokens are the projection's own bias b (they are zero AFTER the norm, image_encoder.py:168-172)."""
    from ullsam_b200 import _lib, ops

    B, heads, D = 2, 12, 768
    M, N = B * 4096, 3 * D
    x, w, b, ref, wf, s, c = _fold_case(M, N, D, 22)
    if partition:
        q = torch.full((B * 25, heads, 196, 64), 7.0, dtype=torch.bfloat16, device="cuda")
        k, v = q.clone(), q.clone()
        ops.gemm(x.bfloat16(), wf, c, epi=_lib.EPI_QKV_FOLD, qkv=(q, k, v), T=196, heads=heads, ln_partition=True, fold_s=s,
                 pad_bias=b, row_stats=_row_stats(x), ln_eps=1e-6)
        y = ref.reshape(B, 64, 64, N)
        yp = b.expand(B, 70, 70, N).clone()
        yp[:, :64, :64] = y
        r = yp.reshape(B, 5, 14, 5, 14, 3, heads, 64).permute(5, 0, 1, 3, 6, 2, 4, 7).reshape(3, B * 25, heads, 196, 64)
        valid = torch.nn.functional.pad(torch.ones(B, 64, 64, device="cuda"), (0, 6, 0, 6))
        valid = valid.reshape(B, 5, 14, 5, 14).permute(0, 1, 3, 2, 4).reshape(B * 25, 1, 196, 1).bool()
        _close(k, r[1], 1.2e-2, "folded windowed k")
        _close(v, r[2], 1.2e-2, "folded windowed v")
        _close(torch.where(valid, q.float(), r[0]), r[0], 1.2e-2, "folded windowed q (real tokens)")
        assert (q.float()[(~valid).expand_as(q)] == 7.0).all()
    else:
        T, G = 4096, B
        q = torch.zeros(G, heads, T, 64, dtype=torch.bfloat16, device="cuda")
        k, v = torch.zeros_like(q), torch.zeros_like(q)
        ops.gemm(x.bfloat16(), wf, c, epi=_lib.EPI_QKV_FOLD, qkv=(q, k, v), T=T, heads=heads, fold_s=s, row_stats=_row_stats(x),
                 ln_eps=1e-6)
        r = ref.reshape(G, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
        for name, got, rr in (("q", q, r[0]), ("k", k, r[1]), ("v", v, r[2])):
            _close(got, rr, 1.2e-2, f"folded qkv[{name}]")


@pytest.mark.parametrize("M,act", [(2 * 4096, 0), (4096 + 77, 2)])
def test_gemm_bf16_with_position_addend(M, act):
    """The decoder's image-side projections (transformer.py:166-167, 178-179): (keys + key_pe) W^T + b is computed as
    keys W^T + b + (key_pe W^T), the second term being one fp32 [4096, N] table shared by all prompts, added in the epilogue
    (after the transpose, with coalesced loads) before the single rounding to bf16."""
    from ullsam_b200 import _lib, ops

    a, w, b = _mk(M, 384, 256, 6)
    pos = torch.randn(4096, 384, device="cuda") * 0.3
    out = ops.gemm(a, w, b, epi=_lib.EPI_BF16, act=act, pos=pos)
    ref = _ref(a, w, b) + pos.repeat((M + 4095) // 4096, 1)[:M]
    if act == 2:
        ref = torch.relu(ref)
    _close(out, ref, 8e-3, f"bf16 + pos, M={M} act={act}")


@pytest.mark.parametrize("M", [1024, 4096 + 50])
def test_gemm_residual_layernorm_256(M):
    """TwoWayAttentionBlock's image side: keys = norm4(keys + out_proj(attn)) (transformer.py:178-182), in place on the
    fp32 keys with a bf16 copy for the next GEMM."""
    from ullsam_b200 import _lib, ops

    a, w, b = _mk(M, 256, 128, 7)
    g = torch.Generator(device="cuda").manual_seed(8)
    keys = torch.randn(M, 256, device="cuda", generator=g)
    ln_w = torch.randn(256, device="cuda", generator=g) * 0.1 + 1
    ln_b = torch.randn(256, device="cuda", generator=g) * 0.1
    ref = torch.nn.functional.layer_norm(_ref(a, w, b) + keys, (256,), ln_w, ln_b, 1e-5)
    x = keys.clone()
    out2 = torch.empty(M, 256, dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, w, b, epi=_lib.EPI_LN_RES, out=x, residual=x, ln_w=ln_w, ln_b=ln_b, ln_eps=1e-5, out2=out2)   # in place
    _close(x, ref, 2e-3, f"LN(res + gemm) fp32, M={M}")
    _close(out2, ref, 8e-3, f"LN(res + gemm) bf16 copy, M={M}")


def _ln2d(x, w, b, eps):
    u = x.mean(-1, keepdim=True)
    s = (x - u).pow(2).mean(-1, keepdim=True)
    return (x - u) / torch.sqrt(s + eps) * w + b


def test_gemm_neck_ln_epilogues():
    from ullsam_b200 import _lib, ops

    B = 2
    a, w, _ = _mk(B * 4096, 256, 768, 5)
    lw = torch.randn(256, device="cuda") * 0.2 + 1.0
    lb = torch.randn(256, device="cuda") * 0.1
    ref = _ln2d(a.float() @ w.float().t(), lw, lb, 1e-6)
    out = ops.gemm(a, w, None, epi=_lib.EPI_LN_BF16, ln_w=lw, ln_b=lb, ln_eps=1e-6)
    _close(out, ref, 1e-2, "ln_bf16")
    out2 = ops.gemm(a, w, None, epi=_lib.EPI_LN_NCHW, ln_w=lw, ln_b=lb, ln_eps=1e-6, hw=4096)
    _close(out2, ref.reshape(B, 4096, 256).permute(0, 2, 1), 2e-4, "ln_nchw")


def test_gemm_conv3x3_implicit():
    from ullsam_b200 import _lib, ops

    B, Cc = 2, 256
    g = torch.Generator(device="cuda").manual_seed(6)
    x = (torch.randn(B, 64, 64, Cc, device="cuda", generator=g) * 0.5).bfloat16()
    wt = (torch.randn(256, Cc, 3, 3, device="cuda", generator=g) * 0.03).bfloat16()
    lw = torch.ones(256, device="cuda")
    lb = torch.zeros(256, device="cuda")
    # kernel weight layout: [out][ (ky*3+kx)*C + c ]
    wk = wt.permute(0, 2, 3, 1).reshape(256, 9 * Cc).contiguous()
    conv = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1)  # [B,256,64,64]
    ref = _ln2d(conv.permute(0, 2, 3, 1), lw, lb, 1e-6).permute(0, 3, 1, 2).reshape(B, 256, 4096)
    out = ops.gemm(x, wk, None, epi=_lib.EPI_LN_NCHW, ln_w=lw, ln_b=lb, ln_eps=1e-6, hw=4096, conv3=True)
    _close(out, ref, 2e-4, "conv3x3+ln")
