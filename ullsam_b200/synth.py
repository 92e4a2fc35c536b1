"""Seeded synthetic weights and inputs for the SAM path (there is no network for checkpoints or datasets).

Weights: a ``state_dict`` with exactly the reference's key names and shapes, drawn from a
seeded CPU generator: Linear/Conv weights and biases ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (the magnitude of
PyTorch's default init), LayerNorm affine near identity, and — unlike the reference's zero init
(image_encoder.py:68-70, 221-222), which would leave those code paths untested — pos_embed and
rel_pos_h/rel_pos_w ~ N(0, 0.02).  There is no network access for real checkpoints.

Inputs: "microscopy-shaped" images (a field of Gaussian blobs plus noise, min-max scaled to [0,1], replicated
to 3 channels — the result of the reference's Resize -> ToTensor -> Normalize(0,1) pipeline,
train_joint_v2.py:271-275), a box prompt around one blob per image, and point prompts.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from .config import SamConfig


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound


def _linear(sd, gen, name, out_f, in_f, bias=True):
    b = 1.0 / math.sqrt(in_f)
    sd[name + ".weight"] = _uniform(gen, (out_f, in_f), b)
    if bias:
        sd[name + ".bias"] = _uniform(gen, (out_f,), b)


def _norm(sd, gen, name, dim):
    sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(dim, generator=gen)
    sd[name + ".bias"] = 0.05 * torch.randn(dim, generator=gen)


def make_state_dict(cfg: SamConfig, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Full Sam state_dict (image_encoder.*, prompt_encoder.*, mask_decoder.*) for ``cfg``."""
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    D, g, hd = cfg.embed_dim, cfg.grid, cfg.embed_dim // cfg.num_heads
    p = "image_encoder."
    sd[p + "pos_embed"] = 0.02 * torch.randn(1, g, g, D, generator=gen)
    fan = cfg.in_chans * cfg.patch_size ** 2
    sd[p + "patch_embed.proj.weight"] = _uniform(gen, (D, cfg.in_chans, cfg.patch_size, cfg.patch_size), 1 / math.sqrt(fan))
    sd[p + "patch_embed.proj.bias"] = _uniform(gen, (D,), 1 / math.sqrt(fan))
    mlp = int(D * cfg.mlp_ratio)
    for i in range(cfg.depth):
        b = f"{p}blocks.{i}."
        size = g if i in cfg.global_attn_indexes else cfg.window_size
        _norm(sd, gen, b + "norm1", D)
        sd[b + "attn.rel_pos_h"] = 0.02 * torch.randn(2 * size - 1, hd, generator=gen)
        sd[b + "attn.rel_pos_w"] = 0.02 * torch.randn(2 * size - 1, hd, generator=gen)
        _linear(sd, gen, b + "attn.qkv", 3 * D, D)
        _linear(sd, gen, b + "attn.proj", D, D)
        _norm(sd, gen, b + "norm2", D)
        _linear(sd, gen, b + "mlp.lin1", mlp, D)
        _linear(sd, gen, b + "mlp.lin2", D, mlp)
    C = cfg.out_chans
    sd[p + "neck.0.weight"] = _uniform(gen, (C, D, 1, 1), 1 / math.sqrt(D))
    _norm(sd, gen, p + "neck.1", C)
    sd[p + "neck.2.weight"] = _uniform(gen, (C, C, 3, 3), 1 / math.sqrt(9 * C))
    _norm(sd, gen, p + "neck.3", C)

    # ---- prompt encoder (prompt_encoder.py:40-63, 215-218) ----
    q = "prompt_encoder."
    E = cfg.transformer_dim
    sd[q + "pe_layer.positional_encoding_gaussian_matrix"] = torch.randn(2, E // 2, generator=gen)
    for i in range(4):
        sd[f"{q}point_embeddings.{i}.weight"] = torch.randn(1, E, generator=gen)
    sd[q + "not_a_point_embed.weight"] = torch.randn(1, E, generator=gen)
    sd[q + "no_mask_embed.weight"] = torch.randn(1, E, generator=gen)
    sd[q + "llm_scale_factor"] = torch.full((1,), 0.1)
    sd[q + "llm_bias"] = torch.zeros(1)
    m = cfg.mask_in_chans
    sd[q + "mask_downscaling.0.weight"] = _uniform(gen, (m // 4, 1, 2, 2), 0.5)
    sd[q + "mask_downscaling.0.bias"] = _uniform(gen, (m // 4,), 0.5)
    _norm(sd, gen, q + "mask_downscaling.1", m // 4)
    sd[q + "mask_downscaling.3.weight"] = _uniform(gen, (m, m // 4, 2, 2), 0.25)
    sd[q + "mask_downscaling.3.bias"] = _uniform(gen, (m,), 0.25)
    _norm(sd, gen, q + "mask_downscaling.4", m)
    sd[q + "mask_downscaling.6.weight"] = _uniform(gen, (E, m, 1, 1), 0.25)
    sd[q + "mask_downscaling.6.bias"] = _uniform(gen, (E,), 0.25)

    # ---- mask decoder (mask_decoder.py:43-69, transformer.py:45-60, 130-151) ----
    d = "mask_decoder."
    t = d + "transformer."

    def attn(name, internal):
        for proj in ("q_proj", "k_proj", "v_proj"):
            _linear(sd, gen, f"{name}.{proj}", internal, E)
        _linear(sd, gen, f"{name}.out_proj", E, internal)

    for i in range(cfg.dec_depth):
        L = f"{t}layers.{i}."
        attn(L + "self_attn", E)
        _norm(sd, gen, L + "norm1", E)
        attn(L + "cross_attn_token_to_image", E // 2)
        _norm(sd, gen, L + "norm2", E)
        _linear(sd, gen, L + "mlp.lin1", cfg.dec_mlp_dim, E)
        _linear(sd, gen, L + "mlp.lin2", E, cfg.dec_mlp_dim)
        _norm(sd, gen, L + "norm3", E)
        _norm(sd, gen, L + "norm4", E)
        attn(L + "cross_attn_image_to_token", E // 2)
    attn(t + "final_attn_token_to_image", E // 2)
    _norm(sd, gen, t + "norm_final_attn", E)
    sd[d + "iou_token.weight"] = torch.randn(1, E, generator=gen)
    sd[d + "mask_tokens.weight"] = torch.randn(cfg.num_mask_tokens, E, generator=gen)
    sd[d + "output_upscaling.0.weight"] = _uniform(gen, (E, E // 4, 2, 2), 1 / math.sqrt(E))
    sd[d + "output_upscaling.0.bias"] = _uniform(gen, (E // 4,), 1 / math.sqrt(E))
    _norm(sd, gen, d + "output_upscaling.1", E // 4)
    sd[d + "output_upscaling.3.weight"] = _uniform(gen, (E // 4, E // 8, 2, 2), 1 / math.sqrt(E // 4))
    sd[d + "output_upscaling.3.bias"] = _uniform(gen, (E // 8,), 1 / math.sqrt(E // 4))
    for i in range(cfg.num_mask_tokens):
        h = f"{d}output_hypernetworks_mlps.{i}."
        _linear(sd, gen, h + "layers.0", E, E)
        _linear(sd, gen, h + "layers.1", E, E)
        _linear(sd, gen, h + "layers.2", E // 8, E)
    h = d + "iou_prediction_head."
    dims = [E] + [cfg.iou_head_hidden_dim] * (cfg.iou_head_depth - 1) + [cfg.num_mask_tokens]
    for i in range(cfg.iou_head_depth):
        _linear(sd, gen, f"{h}layers.{i}", dims[i + 1], dims[i])
    return sd


def make_images(B: int, size: int = 1024, seed: int = 7) -> Tuple[torch.Tensor, torch.Tensor]:
    """[B,3,size,size] fp32 in [0,1] plus one XYXY box per image around a blob ([B,4], input-frame pixels)."""
    gen = torch.Generator().manual_seed(seed)
    ys = torch.arange(size, dtype=torch.float32)[:, None]
    xs = torch.arange(size, dtype=torch.float32)[None, :]
    imgs, boxes = [], []
    for _ in range(B):
        n = int(torch.randint(50, 201, (1,), generator=gen))
        cy = torch.rand(n, generator=gen) * size
        cx = torch.rand(n, generator=gen) * size
        sg = (8.0 + 32.0 * torch.rand(n, generator=gen)) * size / 1024.0
        amp = 0.3 + 0.7 * torch.rand(n, generator=gen)
        field = torch.zeros(size, size)
        for i in range(n):
            # separable Gaussian: outer product of two 1-D profiles
            gy = torch.exp(-0.5 * ((ys[:, 0] - cy[i]) / sg[i]) ** 2)
            gx = torch.exp(-0.5 * ((xs[0] - cx[i]) / sg[i]) ** 2)
            field += amp[i] * gy[:, None] * gx[None, :]
        field = field + 0.05 * torch.randn(size, size, generator=gen)
        field = (field - field.min()) / (field.max() - field.min())
        imgs.append(field[None].expand(3, -1, -1))
        r = 2.0 * sg[0]
        boxes.append(torch.stack([cx[0] - r, cy[0] - r, cx[0] + r, cy[0] + r]).clamp(0, size - 1))
    return torch.stack(imgs).contiguous(), torch.stack(boxes)


def make_points(B: int, size: int = 1024, seed: int = 9) -> Tuple[torch.Tensor, torch.Tensor]:
    """uLLSAM-style point prompts: 1 positive + 3 negative points per instance (train_joint_v2.py:1033-1041)."""
    gen = torch.Generator().manual_seed(seed)
    coords = torch.rand(B, 4, 2, generator=gen) * (size - 1)
    labels = torch.tensor([[1, 0, 0, 0]]).expand(B, -1).clone()
    return coords, labels


def make_llm_map(P: int, seed: int = 11, C: int = 256, g: int = 64) -> torch.Tensor:
    """Stand-in for InternVLSAMModel.text_aware_dense_feature's [P,256,64,64] map (modeling_internvl_sam.py:253-270)."""
    gen = torch.Generator().manual_seed(seed)
    return torch.randn(P, C, g, g, generator=gen)


def make_prompt_embeddings(B: int, Ns: int = 2, seed: int = 13, C: int = 256, g: int = 64):
    """Synthetic stand-ins for PromptEncoder's outputs (the path's inputs; the module itself stays in PyTorch):
    sparse [B,Ns,C], image_pe [1,C,g,g] (a random-Fourier positional encoding) and the no-mask dense vector [C]."""
    gen = torch.Generator().manual_seed(seed)
    sparse = torch.randn(B, Ns, C, generator=gen)
    gauss = torch.randn(2, C // 2, generator=gen)
    ys = (torch.arange(g, dtype=torch.float32) + 0.5) / g
    coords = torch.stack([ys[None, :].expand(g, g), ys[:, None].expand(g, g)], dim=-1)
    proj = 2.0 * math.pi * ((2.0 * coords - 1.0) @ gauss)
    pe = torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1).permute(2, 0, 1)[None].contiguous()
    dense_vec = torch.randn(C, generator=gen)
    return sparse, pe, dense_vec
