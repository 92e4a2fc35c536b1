"""CPU oracle for the SAM segmentation branch of uLLSAM — TEST INFRASTRUCTURE, NOT A PRODUCT PATH.

A functional fp32 restatement (plain torch CPU ops over a ``state_dict``) of the algorithm in the
reference's

    modeling/image_encoder.py   ImageEncoderViT / Block / Attention / window_* / add_decomposed_rel_pos / PatchEmbed
    modeling/common.py          MLPBlock, LayerNorm2d
    modeling/transformer.py     TwoWayTransformer / TwoWayAttentionBlock / Attention
    modeling/mask_decoder.py    MaskDecoder / MLP
    modeling/prompt_encoder.py  PromptEncoder (only to manufacture the path's inputs)

Each function cites the reference file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module, and only as the checker or
the timed CPU baseline; nothing under ``ullsam_b200/`` imports it.

Pinning: the reference ships no tests or golden vectors.  ``oracle/make_golden.py`` runs the
reference's own modules (imported from /root/reference in the build container) on the seeded synthetic weights
and inputs of ``ullsam_b200/synth.py`` and commits their outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from ullsam_b200.config import SamConfig, config  # noqa: F401  (shared hyper-parameter table)

SD = Dict[str, torch.Tensor]


# =====================================================================================================
# encoder
# =====================================================================================================
def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def layer_norm_2d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """common.py:31-43 — channel LayerNorm on NCHW, biased variance, eps inside the sqrt."""
    mu = x.mean(dim=1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=1, keepdim=True)
    y = (x - mu) / torch.sqrt(var + eps)
    return y * w[None, :, None, None] + b[None, :, None, None]


def gelu(x: torch.Tensor) -> torch.Tensor:
    """nn.GELU() default = exact erf form (common.py:18,23)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def patch_embed(sd: SD, x: torch.Tensor, cfg: SamConfig, prefix: str = "image_encoder.") -> torch.Tensor:
    """image_encoder.py:387-395 + 107-109: k=stride=patch conv == per-patch matmul; NCHW -> NHWC; + pos_embed."""
    w = sd[prefix + "patch_embed.proj.weight"]  # [D, C, p, p]
    b = sd[prefix + "patch_embed.proj.bias"]
    B, Cc, H, W = x.shape
    p = cfg.patch_size
    gh, gw = H // p, W // p
    patches = x.reshape(B, Cc, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(B, gh, gw, Cc * p * p)
    tok = patches @ w.reshape(w.shape[0], -1).t() + b
    pos = sd.get(prefix + "pos_embed")
    if pos is not None:
        tok = tok + pos
    return tok


def window_partition(x: torch.Tensor, ws: int):
    """image_encoder.py:243-264 — zero-pad bottom/right to a multiple of ws, split into ws x ws windows."""
    B, H, W, Cc = x.shape
    ph, pw = (-H) % ws, (-W) % ws
    Hp, Wp = H + ph, W + pw
    xp = x.new_zeros(B, Hp, Wp, Cc)
    xp[:, :H, :W] = x
    win = xp.reshape(B, Hp // ws, ws, Wp // ws, ws, Cc).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, Cc)
    return win, (Hp, Wp)


def window_unpartition(win: torch.Tensor, ws: int, pad_hw, hw) -> torch.Tensor:
    """image_encoder.py:267-289 — inverse of window_partition, cropping the padding."""
    Hp, Wp = pad_hw
    H, W = hw
    nh, nw = Hp // ws, Wp // ws
    B = win.shape[0] // (nh * nw)
    x = win.reshape(B, nh, nw, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, -1)
    return x[:, :H, :W]


def rel_pos_table(size: int, rel_pos: torch.Tensor) -> torch.Tensor:
    """image_encoder.py:292-322 for q_size == k_size == size: R[i, j] = rel_pos[i - j + size - 1].
    (The interpolation branch at :304-313 is never taken on this path: the parameter already has
    2*size-1 rows)"""
    assert rel_pos.shape[0] == 2 * size - 1, "rel_pos interpolation is outside the hot path"
    idx = torch.arange(size)[:, None] - torch.arange(size)[None, :] + (size - 1)
    return rel_pos[idx]  # [size, size, C]


def encoder_attention(sd: SD, x: torch.Tensor, pfx: str, heads: int) -> torch.Tensor:
    """image_encoder.py:224-240 with add_decomposed_rel_pos (:325-361).  x: [Bw, H, W, D]."""
    Bw, H, W, D = x.shape
    hd = D // heads
    qkv = x.reshape(Bw, H * W, D) @ sd[pfx + "qkv.weight"].t() + sd[pfx + "qkv.bias"]
    qkv = qkv.reshape(Bw, H * W, 3, heads, hd).permute(2, 0, 3, 1, 4)  # [3, Bw, heads, T, hd]
    q, k, v = qkv[0], qkv[1], qkv[2]
    scores = (q * hd ** -0.5) @ k.transpose(-1, -2)  # [Bw, heads, T, T]
    Rh = rel_pos_table(H, sd[pfx + "rel_pos_h"])
    Rw = rel_pos_table(W, sd[pfx + "rel_pos_w"])
    q5 = q.reshape(Bw, heads, H, W, hd)  # bias uses the UNSCALED q (:234)
    rel_h = torch.einsum("bnhwc,hkc->bnhwk", q5, Rh)
    rel_w = torch.einsum("bnhwc,wkc->bnhwk", q5, Rw)
    scores = (scores.reshape(Bw, heads, H, W, H, W) + rel_h[..., :, None] + rel_w[..., None, :]).reshape(
        Bw, heads, H * W, H * W)
    attn = torch.softmax(scores, dim=-1)
    o = (attn @ v).permute(0, 2, 1, 3).reshape(Bw, H, W, D)
    return o @ sd[pfx + "proj.weight"].t() + sd[pfx + "proj.bias"]


def encoder_block(sd: SD, x: torch.Tensor, i: int, cfg: SamConfig, prefix: str = "image_encoder.") -> torch.Tensor:
    """image_encoder.py:166-182."""
    p = f"{prefix}blocks.{i}."
    windowed = i not in cfg.global_attn_indexes
    h = layer_norm(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.enc_ln_eps)
    if windowed:
        H, W = h.shape[1], h.shape[2]
        h, pad_hw = window_partition(h, cfg.window_size)
    h = encoder_attention(sd, h, p + "attn.", cfg.num_heads)
    if windowed:
        h = window_unpartition(h, cfg.window_size, pad_hw, (H, W))
    x = x + h
    m = layer_norm(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.enc_ln_eps)
    m = gelu(m @ sd[p + "mlp.lin1.weight"].t() + sd[p + "mlp.lin1.bias"])
    m = m @ sd[p + "mlp.lin2.weight"].t() + sd[p + "mlp.lin2.bias"]
    return x + m


def neck(sd: SD, x: torch.Tensor, prefix: str = "image_encoder.") -> torch.Tensor:
    """image_encoder.py:88-104,114: 1x1 conv -> LayerNorm2d -> 3x3 conv (pad 1) -> LayerNorm2d; NHWC in, NCHW out."""
    y = F.conv2d(x.permute(0, 3, 1, 2), sd[prefix + "neck.0.weight"])
    y = layer_norm_2d(y, sd[prefix + "neck.1.weight"], sd[prefix + "neck.1.bias"])
    y = F.conv2d(y, sd[prefix + "neck.2.weight"], padding=1)
    return layer_norm_2d(y, sd[prefix + "neck.3.weight"], sd[prefix + "neck.3.bias"])


def encoder_forward(sd: SD, images: torch.Tensor, cfg: SamConfig, prefix: str = "image_encoder.",
                    stages: Optional[dict] = None) -> torch.Tensor:
    """ImageEncoderViT.forward, image_encoder.py:106-116.  ``stages`` (optional dict) receives intermediates."""
    x = patch_embed(sd, images.float(), cfg, prefix)
    if stages is not None:
        stages["patch"] = x
    for i in range(cfg.depth):
        x = encoder_block(sd, x, i, cfg, prefix)
        if stages is not None:
            stages[f"block{i}"] = x
    return neck(sd, x, prefix)


# =====================================================================================================
# prompt encoder (produces the decoder's inputs; reference modeling/prompt_encoder.py)
# =====================================================================================================
def _pe_encoding(gauss: torch.Tensor, coords01: torch.Tensor) -> torch.Tensor:
    """prompt_encoder.py:220-228: coords in [0,1]^2 -> [sin, cos](2*pi*(2c-1) @ G)."""
    c = (2.0 * coords01 - 1.0).to(gauss.dtype) @ gauss
    c = 2.0 * math.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)


def dense_pe(sd: SD, cfg: SamConfig, prefix: str = "prompt_encoder.") -> torch.Tensor:
    """PromptEncoder.get_dense_pe, prompt_encoder.py:65-74, 230-241 -> [1, C, g, g]."""
    g = cfg.grid
    gauss = sd[prefix + "pe_layer.positional_encoding_gaussian_matrix"]
    ys = (torch.arange(g, dtype=torch.float32, device=gauss.device) + 0.5) / g
    xs = (torch.arange(g, dtype=torch.float32, device=gauss.device) + 0.5) / g
    coords = torch.stack([xs[None, :].expand(g, g), ys[:, None].expand(g, g)], dim=-1)
    return _pe_encoding(gauss, coords).permute(2, 0, 1)[None]


def embed_boxes(sd: SD, boxes: torch.Tensor, cfg: SamConfig, prefix: str = "prompt_encoder.") -> torch.Tensor:
    """prompt_encoder.py:96-103: XYXY boxes [P,4] -> [P,2,C] corner embeddings."""
    gauss = sd[prefix + "pe_layer.positional_encoding_gaussian_matrix"]
    c = (boxes.float() + 0.5).reshape(-1, 2, 2) / float(cfg.img_size)
    e = _pe_encoding(gauss, c)
    e[:, 0, :] += sd[prefix + "point_embeddings.2.weight"][0]
    e[:, 1, :] += sd[prefix + "point_embeddings.3.weight"][0]
    return e


def embed_points(sd: SD, coords: torch.Tensor, labels: torch.Tensor, cfg: SamConfig, pad: bool = True,
                 prefix: str = "prompt_encoder.") -> torch.Tensor:
    """prompt_encoder.py:76-94: points [P,N,2], labels [P,N] (1 pos, 0 neg, -1 pad) -> [P,N(+1),C]."""
    gauss = sd[prefix + "pe_layer.positional_encoding_gaussian_matrix"]
    pts = coords.float() + 0.5
    lab = labels.clone()
    if pad:
        pts = torch.cat([pts, pts.new_zeros(pts.shape[0], 1, 2)], dim=1)
        lab = torch.cat([lab, -lab.new_ones(lab.shape[0], 1)], dim=1)
    e = _pe_encoding(gauss, pts / float(cfg.img_size))
    e = torch.where((lab == -1)[..., None], sd[prefix + "not_a_point_embed.weight"][0].expand_as(e), e)
    e = e + (lab == 0)[..., None] * sd[prefix + "point_embeddings.0.weight"][0]
    e = e + (lab == 1)[..., None] * sd[prefix + "point_embeddings.1.weight"][0]
    return e


def dense_no_mask(sd: SD, P: int, cfg: SamConfig, prefix: str = "prompt_encoder.") -> torch.Tensor:
    """prompt_encoder.py:196-199: no_mask_embed broadcast to [P, C, g, g]."""
    g = cfg.grid
    return sd[prefix + "no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(P, -1, g, g)


def dense_from_llm(sd: SD, llm_map: torch.Tensor, prefix: str = "prompt_encoder.") -> torch.Tensor:
    """uLLSAM's LLM-conditioned dense prompt, prompt_encoder.py:131-151, 187-193: per-pixel LayerNorm over
    channels (no affine, eps 1e-5) then * llm_scale_factor + llm_bias."""
    mu = llm_map.mean(dim=1, keepdim=True)
    var = ((llm_map - mu) ** 2).mean(dim=1, keepdim=True)
    n = (llm_map - mu) / torch.sqrt(var + 1e-5)
    return n * sd[prefix + "llm_scale_factor"] + sd[prefix + "llm_bias"]


def embed_masks(sd: SD, masks: torch.Tensor, prefix: str = "prompt_encoder.") -> torch.Tensor:
    """Stock SAM's mask-prompt branch, prompt_encoder.py:53-61, 105-108, 176-177 (mask_downscaling):
    Conv2d(1, c/4, k2, s2) -> LayerNorm2d -> GELU -> Conv2d(c/4, c, k2, s2) -> LayerNorm2d -> GELU -> Conv2d(c, E, 1).
    masks [P, 1, 256, 256] -> [P, 256, 64, 64]."""
    p = prefix + "mask_downscaling."
    x = F.conv2d(masks, sd[p + "0.weight"], sd[p + "0.bias"], stride=2)
    x = gelu(layer_norm_2d(x, sd[p + "1.weight"], sd[p + "1.bias"]))
    x = F.conv2d(x, sd[p + "3.weight"], sd[p + "3.bias"], stride=2)
    x = gelu(layer_norm_2d(x, sd[p + "4.weight"], sd[p + "4.bias"]))
    return F.conv2d(x, sd[p + "6.weight"], sd[p + "6.bias"])


# =====================================================================================================
# decoder
# =====================================================================================================
def _dec_attention(sd: SD, pfx: str, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """transformer.py:220-242: separate projections, split heads, QK^T / sqrt(c_per_head), softmax, out_proj."""
    q = q @ sd[pfx + "q_proj.weight"].t() + sd[pfx + "q_proj.bias"]
    k = k @ sd[pfx + "k_proj.weight"].t() + sd[pfx + "k_proj.bias"]
    v = v @ sd[pfx + "v_proj.weight"].t() + sd[pfx + "v_proj.bias"]

    def split(t):
        b, n, c = t.shape
        return t.reshape(b, n, heads, c // heads).transpose(1, 2)

    q, k, v = split(q), split(k), split(v)
    a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1]), dim=-1)
    o = (a @ v).transpose(1, 2)
    o = o.reshape(o.shape[0], o.shape[1], -1)
    return o @ sd[pfx + "out_proj.weight"].t() + sd[pfx + "out_proj.bias"]


def _ln(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return layer_norm(x, sd[name + ".weight"], sd[name + ".bias"], 1e-5)  # nn.LayerNorm default eps


def transformer_forward(sd: SD, image_embedding: torch.Tensor, image_pe: torch.Tensor, point_embedding: torch.Tensor,
                        cfg: SamConfig, prefix: str = "mask_decoder.transformer.",
                        skip_first_layer_pe: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """TwoWayTransformer.forward, transformer.py:62-108 (blocks: :153-184).  ``skip_first_layer_pe`` is the constructor flag of
    TwoWayAttentionBlock (:134, set to ``i == 0`` at :52); False is only used by the tests as a negative control."""
    keys = image_embedding.flatten(2).permute(0, 2, 1)
    key_pe = image_pe.flatten(2).permute(0, 2, 1)
    queries = point_embedding
    H = cfg.dec_heads
    for i in range(cfg.dec_depth):
        p = f"{prefix}layers.{i}."
        if i == 0 and skip_first_layer_pe:  # the output REPLACES the queries (:157-158)
            queries = _dec_attention(sd, p + "self_attn.", queries, queries, queries, H)
        else:
            qq = queries + point_embedding
            queries = queries + _dec_attention(sd, p + "self_attn.", qq, qq, queries, H)
        queries = _ln(sd, p + "norm1", queries)
        queries = queries + _dec_attention(sd, p + "cross_attn_token_to_image.", queries + point_embedding,
                                           keys + key_pe, keys, H)
        queries = _ln(sd, p + "norm2", queries)
        m = torch.relu(queries @ sd[p + "mlp.lin1.weight"].t() + sd[p + "mlp.lin1.bias"])
        queries = queries + (m @ sd[p + "mlp.lin2.weight"].t() + sd[p + "mlp.lin2.bias"])
        queries = _ln(sd, p + "norm3", queries)
        keys = keys + _dec_attention(sd, p + "cross_attn_image_to_token.", keys + key_pe, queries + point_embedding,
                                     queries, H)
        keys = _ln(sd, p + "norm4", keys)
    queries = queries + _dec_attention(sd, prefix + "final_attn_token_to_image.", queries + point_embedding,
                                       keys + key_pe, keys, H)
    queries = _ln(sd, prefix + "norm_final_attn", queries)
    return queries, keys


def _mlp3(sd: SD, pfx: str, x: torch.Tensor, n_layers: int) -> torch.Tensor:
    """mask_decoder.py:154-176: Linear/ReLU stack, no activation after the last layer."""
    for i in range(n_layers):
        x = x @ sd[f"{pfx}layers.{i}.weight"].t() + sd[f"{pfx}layers.{i}.bias"]
        if i < n_layers - 1:
            x = torch.relu(x)
    return x


def conv_transpose_2x2(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d(k=2, s=2): out[:, co, 2i+a, 2j+c] = sum_ci x[:, ci, i, j] * w[ci, co, a, c] + b[co]."""
    B, Ci, H, W = x.shape
    Co = w.shape[1]
    y = torch.einsum("bihw,ioac->bohawc", x, w).reshape(B, Co, 2 * H, 2 * W)
    return y + b[None, :, None, None]


def predict_masks(sd: SD, image_embeddings: torch.Tensor, image_pe: torch.Tensor, sparse: torch.Tensor,
                  dense: torch.Tensor, cfg: SamConfig, prefix: str = "mask_decoder.",
                  stages: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """MaskDecoder.predict_masks, mask_decoder.py:112-149."""
    P = sparse.shape[0]
    nm = cfg.num_mask_tokens
    out_tok = torch.cat([sd[prefix + "iou_token.weight"], sd[prefix + "mask_tokens.weight"]], dim=0)
    tokens = torch.cat([out_tok[None].expand(P, -1, -1), sparse], dim=1)
    src = torch.repeat_interleave(image_embeddings, P, dim=0) + dense
    pos = torch.repeat_interleave(image_pe, P, dim=0)
    b, c, h, w = src.shape
    hs, keys = transformer_forward(sd, src, pos, tokens, cfg, prefix + "transformer.")
    iou_tok = hs[:, 0]
    mask_tok = hs[:, 1:1 + nm]
    up = keys.transpose(1, 2).reshape(b, c, h, w)
    up = conv_transpose_2x2(up, sd[prefix + "output_upscaling.0.weight"], sd[prefix + "output_upscaling.0.bias"])
    up = gelu(layer_norm_2d(up, sd[prefix + "output_upscaling.1.weight"], sd[prefix + "output_upscaling.1.bias"]))
    up = gelu(conv_transpose_2x2(up, sd[prefix + "output_upscaling.3.weight"], sd[prefix + "output_upscaling.3.bias"]))
    hyper = torch.stack(
        [_mlp3(sd, f"{prefix}output_hypernetworks_mlps.{i}.", mask_tok[:, i], 3) for i in range(nm)], dim=1)
    bb, cc, hh, ww = up.shape
    masks = (hyper @ up.reshape(bb, cc, hh * ww)).reshape(bb, nm, hh, ww)
    iou = _mlp3(sd, prefix + "iou_prediction_head.", iou_tok, cfg.iou_head_depth)
    if stages is not None:
        stages.update(queries=hs, keys=keys, upscaled=up, hyper_in=hyper)
    return masks, iou


def decoder_forward(sd: SD, image_embeddings: torch.Tensor, image_pe: torch.Tensor, sparse: torch.Tensor,
                    dense: torch.Tensor, multimask_output: bool, cfg: SamConfig,
                    prefix: str = "mask_decoder.") -> Tuple[torch.Tensor, torch.Tensor]:
    """MaskDecoder.forward, mask_decoder.py:71-109: slice 0 for a single mask, 1: for multimask."""
    masks, iou = predict_masks(sd, image_embeddings, image_pe, sparse, dense, cfg, prefix)
    sl = slice(1, None) if multimask_output else slice(0, 1)
    return masks[:, sl], iou[:, sl]


def segment(sd: SD, images: torch.Tensor, sparse: torch.Tensor, dense: torch.Tensor, multimask_output: bool,
            cfg: SamConfig):
    """Encoder + per-image decoder calls (the loop of Sam.forward, sam.py:98-117): image i with prompt set i."""
    emb = encoder_forward(sd, images, cfg)
    pe = dense_pe(sd, cfg)
    masks, ious = [], []
    for i in range(images.shape[0]):
        d = dense[i:i + 1] if dense.shape[0] == images.shape[0] else dense
        m, s = decoder_forward(sd, emb[i:i + 1], pe, sparse[i:i + 1], d, multimask_output, cfg)
        masks.append(m)
        ious.append(s)
    return emb, torch.cat(masks), torch.cat(ious)


# =====================================================================================================
# post-processing (downstream of the decoder; reference modeling/sam.py, train_joint_v2.py)
# =====================================================================================================
def postprocess_masks(masks: torch.Tensor, input_size, original_size, img_size: int = 1024) -> torch.Tensor:
    """Sam.postprocess_masks, sam.py:133-162."""
    m = F.interpolate(masks, (img_size, img_size), mode="bilinear", align_corners=False)
    m = m[..., : input_size[0], : input_size[1]]
    return F.interpolate(m, tuple(original_size), mode="bilinear", align_corners=False)


def upsample_masks(masks: torch.Tensor, img_size: int = 1024) -> torch.Tensor:
    """uLLSAM's upsampling of the low-resolution logits to the model size (train_joint_v2.py:1073-1079, app.py:635-641)."""
    return F.interpolate(masks, (img_size, img_size), mode="bilinear", align_corners=False)


def calc_iou(logits: torch.Tensor, targets: torch.Tensor, smooth: float = 1e-7) -> torch.Tensor:
    """CalcIoU.forward, train_joint_v2.py:666-696: [B, N, H, W] x2 -> [B, N]."""
    p = (logits.sigmoid() > 0.5).float().flatten(2)
    t = targets.flatten(2)
    inter = (p * t).sum(-1)
    union = p.sum(-1) + t.sum(-1) - inter
    return (inter + smooth) / (union + smooth)


# =====================================================================================================
# MLLM bridge (either side of the language model; reference modeling/modeling_internvl_sam.py)
# =====================================================================================================
def pixel_shuffle(x: torch.Tensor, scale_factor: float = 0.5, ps_version: str = "v2") -> torch.Tensor:
    """InternVLSAMModel.pixel_shuffle, modeling_internvl_sam.py:211-227 (x is NHWC)."""
    n, h, w, c = x.size()
    x = x.reshape(n, h, int(w * scale_factor), int(c / scale_factor))
    x = x.permute(0, 2, 1, 3).contiguous()
    x = x.reshape(n, int(w * scale_factor), int(h * scale_factor), int(c / (scale_factor * scale_factor)))
    if ps_version != "v1":
        x = x.permute(0, 2, 1, 3).contiguous()
    return x


def _projector(x: torch.Tensor, p: Dict[str, torch.Tensor]) -> torch.Tensor:
    """Sequential(LayerNorm, Linear, GELU, Linear), modeling_internvl_sam.py:88-101 (LayerNorm eps default 1e-5)."""
    x = F.layer_norm(x, (x.shape[-1],), p["0.weight"], p["0.bias"], 1e-5)
    x = F.gelu(F.linear(x, p["1.weight"], p["1.bias"]))
    return F.linear(x, p["3.weight"], p["3.bias"])


def mllm_extract_feature(vit_features: torch.Tensor, mlp1: Dict[str, torch.Tensor], ps_version: str = "v2") -> torch.Tensor:
    """extract_feature after the encoder, modeling_internvl_sam.py:241-251: [B,256,64,64] -> [B,1024,H]."""
    f = pixel_shuffle(vit_features.permute(0, 2, 3, 1), 0.5, ps_version)
    f = f.reshape(f.shape[0], -1, f.shape[-1])
    return _projector(f, mlp1)


def text_aware_dense_feature(features: torch.Tensor, mlp2: Dict[str, torch.Tensor], ps_version: str = "v2",
                             ratio: float = 0.5) -> torch.Tensor:
    """modeling_internvl_sam.py:253-270: [B,1024,H] -> [B,256,64,64]."""
    f = _projector(features, mlp2)
    s = int(math.sqrt(f.shape[1]))
    f = f.reshape(f.shape[0], s, s, f.shape[2])
    if ps_version != "v1":
        f = f.permute(0, 2, 1, 3).contiguous()
    n, h, w, c = f.size()
    f = f.reshape(n, h, int(w / ratio), int(c * ratio))
    f = f.permute(0, 2, 1, 3).contiguous()
    f = f.reshape(n, int(w / ratio), int(h / ratio), int(c * (ratio * ratio)))
    return f.permute(0, 3, 1, 2).contiguous()
