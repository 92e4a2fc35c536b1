"""Operator-level wrappers over the C ABI (``ullsam_op_*``) taking CUDA torch tensors.

These are what the kernel parity tests call; the nn.Module shims in ``ullsam_b200.modeling`` use the
model-level entry points instead.  torch is only the allocator / stream provider here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GemmDesc, check, current_stream_ptr, ptr


import functools


def _require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.UllsamError("ullsam_b200 ops need CUDA tensors; there is no CPU path")


def _on_tensor_device(fn):
    """The op-level ABI launches on the CURRENT device and takes only the stream from the tensor's device: make the
    device of the first tensor argument current for the duration of the call (tensors on a non-current GPU would
    otherwise fail with an invalid resource handle), and restore the caller's device afterwards."""

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        first = next((a for a in list(args) + list(kwargs.values()) if isinstance(a, torch.Tensor)), None)
        if first is None or not first.is_cuda:
            return fn(*args, **kwargs)   # the op raises its own "no CPU path" error
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda and a.device != first.device:
                raise _lib.UllsamError(f"{fn.__name__}: tensors on different devices ({first.device} vs {a.device})")
        with torch.cuda.device(first.device):
            return fn(*args, **kwargs)

    return wrapper


@_on_tensor_device
def gemm(
    a: torch.Tensor,
    w: torch.Tensor,
    bias: Optional[torch.Tensor] = None,
    *,
    epi: int = _lib.EPI_BF16,
    act: int = _lib.ACT_NONE,
    out: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    pos: Optional[torch.Tensor] = None,
    qkv: Optional[tuple] = None,
    T: int = 0,
    heads: int = 0,
    ln_w: Optional[torch.Tensor] = None,
    ln_b: Optional[torch.Tensor] = None,
    ln_eps: float = 1e-6,
    hw: int = 4096,
    conv3: bool = False,
    out2: Optional[torch.Tensor] = None,
    ln_partition: bool = False,
    row_stats: Optional[torch.Tensor] = None,
    fold_s: Optional[torch.Tensor] = None,
    pad_bias: Optional[torch.Tensor] = None,
    kmax_sq: Optional[torch.Tensor] = None,
) -> Optional[torch.Tensor]:
    """C = A[M,K] @ W[N,K]^T with the epilogue selected by ``epi`` (see include/ullsam.h)."""
    _require_cuda(a, w, bias, out, residual, pos, ln_w, ln_b, out2, row_stats, fold_s, pad_bias, kmax_sq)
    if row_stats is not None:  # fp32 [M, STAT_SLOTS, 2]: (sum, sum of squares) per row and 128-column slice
        assert row_stats.dtype == torch.float32 and row_stats.is_contiguous() and row_stats.numel() == a.shape[0] * _lib.STAT_SLOTS * 2
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.is_contiguous() and w.is_contiguous()
    N, K = w.shape
    if conv3:
        B, H, Wd, Cc = a.shape
        M = B * H * Wd
        assert K == 9 * Cc
    else:
        M = a.shape[0]
        assert a.shape[1] == K
    d = GemmDesc()
    d.epi, d.M, d.N, d.K = epi, M, N, K
    d.a_conv3, d.conv_c = (1, a.shape[-1]) if conv3 else (0, 0)
    d.A, d.W, d.bias = ptr(a), ptr(w), ptr(bias)
    d.act = act
    d.ldc = N
    if epi in (_lib.EPI_QKV, _lib.EPI_QKV_FOLD):
        q, k, v = qkv
        d.q, d.k, d.v, d.T, d.heads = ptr(q), ptr(k), ptr(v), T, heads
        # ln_partition: A's rows are image tokens [B,64,64]; the store window-partitions them (pad rows not written)
    else:
        if out is None:
            if epi in (_lib.EPI_BF16, _lib.EPI_LN_BF16, _lib.EPI_GELU_FOLD):
                out = torch.empty((M, N), dtype=torch.bfloat16, device=a.device)
            elif epi == _lib.EPI_LN_NCHW:
                out = torch.empty((M // hw, N, hw), dtype=torch.float32, device=a.device)
            else:
                out = torch.empty((M, N), dtype=torch.float32, device=a.device)
        d.out = ptr(out)
    d.residual = ptr(residual)
    d.pos = ptr(pos)
    d.pos_rows = pos.shape[0] if pos is not None else 0
    d.ln_w, d.ln_b, d.ln_eps, d.hw = ptr(ln_w), ptr(ln_b), ln_eps, hw
    d.out2, d.ln_partition = ptr(out2), int(ln_partition)
    d.row_stats, d.fold_s, d.pad_bias, d.kmax_sq = ptr(row_stats), ptr(fold_s), ptr(pad_bias), ptr(kmax_sq)
    check(_lib.load().ullsam_op_gemm(C.byref(d), current_stream_ptr(a.device)), what="ullsam_op_gemm")
    return out


@_on_tensor_device
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, window: int = 0) -> torch.Tensor:
    """x fp32 [B, grid, grid, dim] -> bf16 LayerNorm; ``window`` > 0 also window-partitions (zero pad rows)."""
    _require_cuda(x, gamma, beta)
    B, g, g2, dim = x.shape
    assert g == g2 and x.dtype == torch.float32 and x.is_contiguous()
    if window > 0:
        nwin = -(-g // window)
        y = torch.empty((B * nwin * nwin, window * window, dim), dtype=torch.bfloat16, device=x.device)
    else:
        y = torch.empty((B, g, g, dim), dtype=torch.bfloat16, device=x.device)
    check(
        _lib.load().ullsam_op_layernorm(ptr(x), ptr(gamma), ptr(beta), eps, B, g, dim, window, ptr(y),
                                        current_stream_ptr(x.device)),
        what="ullsam_op_layernorm",
    )
    return y


def pad_rel_table(rel: torch.Tensor, grid: int) -> torch.Tensor:
    """rel_pos_h / rel_pos_w fp32 [2*grid-1, 64] -> the kernel's bf16 table, zero-padded to 128 / 32 rows."""
    nrel = 128 if grid == 64 else 32
    out = torch.zeros((nrel, 64), dtype=torch.bfloat16, device=rel.device)
    out[: rel.shape[0]] = rel.to(torch.bfloat16)
    return out


@_on_tensor_device
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, rel_h: torch.Tensor, rel_w: torch.Tensor,
              grid: int) -> torch.Tensor:
    """q,k,v bf16 [G, heads, grid*grid, 64]; rel_h/rel_w fp32 [2*grid-1, 64] (the module parameters).
    Returns bf16 [B, 64, 64, heads*64] (windows merged and padding dropped when grid == 14)."""
    _require_cuda(q, k, v, rel_h, rel_w)
    G, heads, T, hd = q.shape
    assert hd == 64 and T == grid * grid and q.dtype == torch.bfloat16
    assert q.is_contiguous() and k.is_contiguous() and v.is_contiguous()
    B = G if grid == 64 else G // 25
    rh, rw = pad_rel_table(rel_h, grid), pad_rel_table(rel_w, grid)
    out = torch.zeros((B, 64, 64, heads * 64), dtype=torch.bfloat16, device=q.device)
    check(
        _lib.load().ullsam_op_attention(ptr(q), ptr(k), ptr(v), ptr(rh), ptr(rw), G, heads, grid, ptr(out),
                                        current_stream_ptr(q.device)),
        what="ullsam_op_attention",
    )
    return out


@_on_tensor_device
def im2col(images: torch.Tensor) -> torch.Tensor:
    _require_cuda(images)
    B = images.shape[0]
    assert images.shape[1:] == (3, 1024, 1024) and images.is_contiguous()
    dt = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}[images.dtype]
    out = torch.empty((B * 4096, 768), dtype=torch.bfloat16, device=images.device)
    check(_lib.load().ullsam_op_im2col(ptr(images), dt, B, ptr(out), current_stream_ptr(images.device)),
          what="ullsam_op_im2col")
    return out


@_on_tensor_device
def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(_lib.load().ullsam_op_cast_bf16(ptr(x), x.numel(), ptr(y), current_stream_ptr(x.device)),
          what="ullsam_op_cast_bf16")
    return y


# ---- PromptEncoder (the step upstream of the decoder) ---------------------------------------------------------------

@_on_tensor_device
def prompt_dense_pe(gauss: torch.Tensor, grid: int) -> torch.Tensor:
    """PositionEmbeddingRandom.forward((grid, grid)) (prompt_encoder.py:230-241): gauss [2, F] -> [2F, grid, grid]."""
    _require_cuda(gauss)
    g = gauss.float().contiguous()
    F = g.shape[1]
    out = torch.empty((2 * F, grid, grid), dtype=torch.float32, device=g.device)
    check(_lib.load().ullsam_op_prompt_dense_pe(ptr(g), F, grid, ptr(out), current_stream_ptr(g.device)),
          what="ullsam_op_prompt_dense_pe")
    return out


@_on_tensor_device
def prompt_sparse(gauss: torch.Tensor, point_emb: torch.Tensor, not_a_point: torch.Tensor,
                  coords: Optional[torch.Tensor], labels: Optional[torch.Tensor], boxes: Optional[torch.Tensor],
                  pad: bool, img_w: float, img_h: float) -> torch.Tensor:
    """_embed_points / _embed_boxes (prompt_encoder.py:76-103): -> [P, Ns, 2F] fp32."""
    _require_cuda(gauss, point_emb, not_a_point, coords, labels, boxes)
    g = gauss.float().contiguous()
    F = g.shape[1]
    pe, nap = point_emb.float().contiguous(), not_a_point.float().contiguous()
    P = coords.shape[0] if coords is not None else boxes.shape[0]
    Np = coords.shape[1] if coords is not None else 0
    c = coords.float().contiguous() if coords is not None else None
    l = labels.float().contiguous() if labels is not None else None
    b = boxes.float().reshape(P, 4).contiguous() if boxes is not None else None
    Ns = (Np + int(pad) if coords is not None else 0) + (2 if boxes is not None else 0)
    out = torch.empty((P, Ns, 2 * F), dtype=torch.float32, device=g.device)
    check(_lib.load().ullsam_op_prompt_sparse(ptr(g), F, ptr(pe), ptr(nap), ptr(c), ptr(l), P, Np, int(pad), ptr(b),
                                              float(img_w), float(img_h), ptr(out), current_stream_ptr(g.device)),
          what="ullsam_op_prompt_sparse")
    return out


@_on_tensor_device
def prompt_llm_dense(x: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """_normalize_llm_hidden_states (prompt_encoder.py:131-151): x [P, C, H, W] fp32/bf16 -> fp32, same shape."""
    _require_cuda(x, scale, bias)
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    x = x.contiguous()
    P, Cc, H, W = x.shape
    dt = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}[x.dtype]
    out = torch.empty((P, Cc, H, W), dtype=torch.float32, device=x.device)
    # the converted copies must outlive the launch: a temporary passed straight to ptr() is freed (and its block handed
    # to the next temporary by the caching allocator) before the kernel reads it
    s = scale.detach().float().contiguous()
    b = bias.detach().float().contiguous()
    check(_lib.load().ullsam_op_prompt_llm_dense(ptr(x), dt, P, Cc, H * W, ptr(s), ptr(b), ptr(out),
                                                 current_stream_ptr(x.device)),
          what="ullsam_op_prompt_llm_dense")
    return out


@_on_tensor_device
def prompt_mask_embed(masks: torch.Tensor, params) -> torch.Tensor:
    """_embed_masks / mask_downscaling (prompt_encoder.py:53-61, 105-108): masks [P, 1, 4G, 4G] -> fp32 [P, E, G, G].
    ``params``: the 10 tensors of ``mask_downscaling`` in state_dict order."""
    _require_cuda(masks, *params)
    if len(params) != 10:
        raise _lib.UllsamError("prompt_mask_embed: expected the 10 tensors of mask_downscaling")
    x = masks.float().contiguous()
    P, one, H, W = x.shape
    if one != 1 or H != W or H % 16 != 0:
        raise _lib.UllsamError(f"prompt_mask_embed: expected [P, 1, 4G, 4G] with G % 4 == 0, got {tuple(masks.shape)}")
    G = H // 4
    ps = [t.detach().float().contiguous() for t in params]     # kept alive until the launch has been queued
    c1, c2, E = ps[0].shape[0], ps[4].shape[0], ps[8].shape[0]
    arr = (C.c_void_p * 10)(*[ptr(t) for t in ps])
    out = torch.empty((P, E, G, G), dtype=torch.float32, device=x.device)
    check(_lib.load().ullsam_op_prompt_mask_embed(ptr(x), P, G, c1, c2, E, arr, ptr(out), current_stream_ptr(x.device)),
          what="ullsam_op_prompt_mask_embed")
    return out


# ---- MLLM bridge ------------------------------------------------------------------------------------------------------

@_on_tensor_device
def pixel_shuffle(x: torch.Tensor, to_tokens: bool, ps_version: str = "v2") -> torch.Tensor:
    """InternVL pixel shuffle with ratio 0.5 (modeling_internvl_sam.py:211-270): emb fp32 [B,256,64,64] -> tokens
    [B,1024,1024] (``to_tokens``) or back."""
    _require_cuda(x)
    x = x.float().contiguous()
    B = x.shape[0]
    if to_tokens:
        assert x.shape[1:] == (256, 64, 64)
        out = torch.empty((B, 1024, 1024), dtype=torch.float32, device=x.device)
    else:
        assert x.shape[1:] == (1024, 1024)
        out = torch.empty((B, 256, 64, 64), dtype=torch.float32, device=x.device)
    check(_lib.load().ullsam_op_pixel_shuffle(ptr(x), B, int(ps_version == "v1"), int(to_tokens), ptr(out),
                                              current_stream_ptr(x.device)), what="ullsam_op_pixel_shuffle")
    return out
