"""Fused attention with decomposed rel-pos bias vs a torch fp32 restatement of image_encoder.py:224-240,325-361.
Tolerance: bf16 P / bf16 output rounding (2^-8 relative) on outputs of magnitude ~|v|."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_attention(q, k, v, rel_h, rel_w, grid):
    """q,k,v: [G, heads, T, 64] fp32 -> [G, heads, T, 64]."""
    G, H, T, hd = q.shape
    idx = torch.arange(grid, device=q.device)[:, None] - torch.arange(grid, device=q.device)[None, :] + grid - 1
    Rh, Rw = rel_h[idx], rel_w[idx]
    out = torch.empty_like(q)
    for g in range(G):
        s = (q[g] * hd ** -0.5) @ k[g].transpose(-1, -2)
        q5 = q[g].reshape(H, grid, grid, hd)
        bh = torch.einsum("nhwc,hkc->nhwk", q5, Rh)
        bw = torch.einsum("nhwc,wkc->nhwk", q5, Rw)
        s = (s.reshape(H, grid, grid, grid, grid) + bh[..., :, None] + bw[..., None, :]).reshape(H, T, T)
        out[g] = torch.softmax(s, dim=-1) @ v[g]
    return out


def _inputs(G, heads, grid, seed, rel_scale):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    T = grid * grid
    q = (torch.randn(G, heads, T, 64, device="cuda", generator=gen) * 1.2).bfloat16()
    k = (torch.randn(G, heads, T, 64, device="cuda", generator=gen) * 1.2).bfloat16()
    v = torch.randn(G, heads, T, 64, device="cuda", generator=gen).bfloat16()
    rel_h = (torch.randn(2 * grid - 1, 64, device="cuda", generator=gen) * rel_scale).bfloat16().float()
    rel_w = (torch.randn(2 * grid - 1, 64, device="cuda", generator=gen) * rel_scale).bfloat16().float()
    return q, k, v, rel_h, rel_w


@pytest.mark.parametrize("rel_scale", [0.0, 0.25])
def test_global_attention(rel_scale):
    from ullsam_b200 import ops

    G, heads, grid = 1, 2, 64
    q, k, v, rel_h, rel_w = _inputs(G, heads, grid, 11, rel_scale)
    out = ops.attention(q, k, v, rel_h, rel_w, grid)  # [B,64,64,heads*64]
    torch.cuda.synchronize()
    ref = _ref_attention(q.float(), k.float(), v.float(), rel_h, rel_w, grid)
    ref = ref.permute(0, 2, 1, 3).reshape(G, 64, 64, heads * 64)
    err = (out.float() - ref).abs().max().item()
    print(f"global rel_scale={rel_scale}: max|err|={err:.3e} ref max={ref.abs().max().item():.3e}")
    assert err < 2e-2


@pytest.mark.parametrize("growth", [6.0, 40.0])
def test_global_attention_moving_row_maximum(growth):
    """Scores that keep growing along the key axis (by `growth` nats in total, and in jumps inside key tiles), so the
    running reference point of the online softmax has to move many times, in the middle of tiles as well as between
    them: exercises the O / row-sum / stored-P rescaling of the fused kernel.  The softmax of every row is dominated by
    the last keys; the reference is the fp32 formula (image_encoder.py:224-240)."""
    from ullsam_b200 import ops

    G, heads, grid = 1, 2, 64
    q, k, v, rel_h, rel_w = _inputs(G, heads, grid, 17, 0.1)
    T = grid * grid
    u = torch.nn.functional.normalize(torch.randn(64, device="cuda"), dim=0)
    ramp = torch.linspace(0.0, 1.0, T, device="cuda")
    steps = (torch.arange(T, device="cuda") // 37).float() / (T // 37)      # jumps that are not aligned with key tiles
    qf = q.float() * 0.3 + 8.0 * u                                         # every query has a large component along u
    kf = k.float() * 0.3 + (growth * (0.5 * ramp + 0.5 * steps))[None, None, :, None] * u
    q, k = qf.bfloat16(), kf.bfloat16()
    out = ops.attention(q, k, v, rel_h, rel_w, grid)
    torch.cuda.synchronize()
    ref = _ref_attention(q.float(), k.float(), v.float(), rel_h, rel_w, grid)
    ref = ref.permute(0, 2, 1, 3).reshape(G, 64, 64, heads * 64)
    err = (out.float() - ref).abs().max().item()
    print(f"global growth={growth}: max|err|={err:.3e} ref max={ref.abs().max().item():.3e}")
    assert torch.isfinite(out.float()).all()
    assert err < 3e-2


@pytest.mark.parametrize("rel_scale", [0.0, 0.25])
def test_windowed_attention(rel_scale):
    from ullsam_b200 import ops

    B, heads, grid = 2, 3, 14
    G = B * 25
    q, k, v, rel_h, rel_w = _inputs(G, heads, grid, 12, rel_scale)
    out = ops.attention(q, k, v, rel_h, rel_w, grid)
    torch.cuda.synchronize()
    ref = _ref_attention(q.float(), k.float(), v.float(), rel_h, rel_w, grid)  # [G, heads, 196, 64]
    # window_unpartition (image_encoder.py:267-289): [B,5,5,14,14,C] -> [B,70,70,C] -> crop 64
    r = ref.permute(0, 2, 1, 3).reshape(B, 5, 5, 14, 14, heads * 64).permute(0, 1, 3, 2, 4, 5).reshape(B, 70, 70, -1)
    r = r[:, :64, :64]
    err = (out.float() - r).abs().max().item()
    print(f"windowed rel_scale={rel_scale}: max|err|={err:.3e} ref max={r.abs().max().item():.3e}")
    assert err < 2e-2


def test_layernorm_and_window_partition():
    from ullsam_b200 import ops

    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(2, 64, 64, 768, device="cuda", generator=gen) * 2 + 0.3
    g = torch.randn(768, device="cuda", generator=gen) * 0.1 + 1
    b = torch.randn(768, device="cuda", generator=gen) * 0.1
    ref = torch.nn.functional.layer_norm(x, (768,), g, b, 1e-6)
    y = ops.layernorm(x, g, b, 1e-6, window=0)
    assert (y.float() - ref).abs().max().item() < 3e-2
    yw = ops.layernorm(x, g, b, 1e-6, window=14)  # [B*25, 196, 768]
    pad = torch.nn.functional.pad(ref, (0, 0, 0, 6, 0, 6))
    refw = pad.reshape(2, 5, 14, 5, 14, 768).permute(0, 1, 3, 2, 4, 5).reshape(50, 196, 768)
    assert (yw.float() - refw).abs().max().item() < 3e-2
    assert (yw[4].reshape(14, 14, 768)[:, 8:] == 0).all()  # right-edge window: columns 64..69 are padding


def test_im2col_matches_unfold():
    from ullsam_b200 import ops

    gen = torch.Generator(device="cuda").manual_seed(4)
    img = torch.rand(2, 3, 1024, 1024, device="cuda", generator=gen)
    ref = img.reshape(2, 3, 64, 16, 64, 16).permute(0, 2, 4, 1, 3, 5).reshape(2 * 4096, 768)
    for t in (img, img.bfloat16()):
        got = ops.im2col(t)
        assert torch.equal(got, ref.to(torch.bfloat16) if t.dtype == torch.float32 else
                           t.reshape(2, 3, 64, 16, 64, 16).permute(0, 2, 4, 1, 3, 5).reshape(2 * 4096, 768))


@pytest.mark.parametrize("grid,G,heads", [(64, 4, 12), (14, 150, 12)])
def test_attention_is_deterministic(grid, G, heads):
    """Race detector: the fused kernel must be bit-reproducible run to run (mbarrier / TMEM hand-offs)."""
    from ullsam_b200 import ops

    q, k, v, rel_h, rel_w = _inputs(G, heads, grid, 21, 0.25)
    ref = ops.attention(q, k, v, rel_h, rel_w, grid)
    for _ in range(4):
        out = ops.attention(q, k, v, rel_h, rel_w, grid)
        assert torch.equal(out, ref)
    assert torch.isfinite(ref.float()).all()
