"""The MLLM bridge of uLLSAM's ``InternVLSAMModel`` on the sm_100a path: the two projector MLPs that
sit either side of the language model, with the pixel shuffle folded into small device kernels.

* :func:`extract_feature`  — ``InternVLSAMModel.extract_feature`` (modeling_internvl_sam.py:241-251):
  SAM embedding [B,256,64,64] -> pixel shuffle (ratio 0.5) -> ``mlp1`` = LayerNorm(1024), Linear(1024,H), GELU, Linear(H,H)
  -> vision tokens [B,1024,H] for the LLM;
* :func:`text_aware_dense_feature` — (:253-270): LLM hidden states [B,1024,H] -> ``mlp2`` = LayerNorm(H),
  Linear(H,1024), GELU, Linear(1024,1024) -> inverse shuffle -> dense prompt map [B,256,64,64].

Both run as: LayerNorm kernel (fp32 statistics) -> tcgen05 GEMM with the exact-erf GELU epilogue -> tcgen05 GEMM, bf16
operands with fp32 accumulation, through ``libullsam_b200.so``.  The callers' second encoder pass
(train_joint_v2.py:1015, app.py:603) is unnecessary: ``extract_feature`` returns the image embeddings it computed
(the reference already exposes them as ``outputs.image_embeddings``), reuse them for the mask decoder.

Usage on a built model (methods are replaced, parameters stay where they are)::

    from ullsam_b200.bridge import patch_internvl_sam
    patch_internvl_sam(model)        # model.extract_feature / model.text_aware_dense_feature -> this module
"""
from __future__ import annotations

import types
from typing import Dict, Tuple

import torch
from torch import nn

from . import _lib, ops

_WCACHE: Dict[Tuple, torch.Tensor] = {}


def _bf16_weight(w: torch.Tensor) -> torch.Tensor:
    """nn.Linear weight [N, K] as the GEMM's bf16 operand, cached per (storage, version)."""
    key = (w.data_ptr(), w._version, w.dtype, tuple(w.shape))
    hit = _WCACHE.get(key)
    if hit is None:
        if len(_WCACHE) > 64:
            _WCACHE.clear()
        hit = w.detach().to(torch.bfloat16).contiguous()
        _WCACHE[key] = hit
    return hit


def _mlp(x_rows: torch.Tensor, mlp: nn.Sequential, out_f32: bool) -> torch.Tensor:
    """LayerNorm -> Linear -> GELU -> Linear on fp32 rows [M, K] (the structure of mlp1 / mlp2, :88-101)."""
    ln, lin1, act, lin2 = mlp[0], mlp[1], mlp[2], mlp[3]
    if not (isinstance(ln, nn.LayerNorm) and isinstance(lin1, nn.Linear) and isinstance(act, nn.GELU)
            and isinstance(lin2, nn.Linear)) or getattr(act, "approximate", "none") != "none":
        raise _lib.UllsamError("bridge: expected Sequential(LayerNorm, Linear, GELU(erf), Linear)")
    M, K = x_rows.shape
    xn = ops.layernorm(x_rows.view(M, 1, 1, K), ln.weight.detach().float().contiguous(), ln.bias.detach().float().contiguous(),
                       ln.eps).view(M, K)
    h = ops.gemm(xn, _bf16_weight(lin1.weight), lin1.bias.detach().float().contiguous(), epi=_lib.EPI_BF16, act=_lib.ACT_GELU)
    return ops.gemm(h, _bf16_weight(lin2.weight), lin2.bias.detach().float().contiguous(),
                    epi=_lib.EPI_F32 if out_f32 else _lib.EPI_BF16)


@torch.no_grad()
def extract_feature(vision_model: nn.Module, mlp1: nn.Sequential, pixel_values: torch.Tensor, ps_version: str = "v2",
                    out_dtype: torch.dtype = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (vision tokens [B, 1024, H], image embeddings [B, 256, 64, 64]) — modeling_internvl_sam.py:241-251."""
    vit_features = vision_model(pixel_values)
    B = vit_features.shape[0]
    tokens = ops.pixel_shuffle(vit_features, to_tokens=True, ps_version=ps_version)   # [B, 1024, 1024] fp32
    out = _mlp(tokens.view(B * 1024, 1024), mlp1, out_f32=True).view(B, 1024, -1)
    return out.to(out_dtype or vit_features.dtype), vit_features


@torch.no_grad()
def text_aware_dense_feature(mlp2: nn.Sequential, features: torch.Tensor, ps_version: str = "v2",
                             out_dtype: torch.dtype = None) -> torch.Tensor:
    """LLM hidden states of the image tokens [B, 1024, H] -> dense map [B, 256, 64, 64] — (:253-270)."""
    B, N, H = features.shape
    if N != 1024:
        raise _lib.UllsamError("bridge: text_aware_dense_feature expects the 1024 image-token positions")
    y = _mlp(features.float().contiguous().view(B * N, H), mlp2, out_f32=True).view(B, 1024, 1024)
    return ops.pixel_shuffle(y, to_tokens=False, ps_version=ps_version).to(out_dtype or features.dtype)


def patch_internvl_sam(model: nn.Module) -> nn.Module:
    """Replace ``extract_feature`` / ``text_aware_dense_feature`` of an ``InternVLSAMModel`` instance
    (attributes used: vision_model, mlp1, mlp2, ps_version, downsample_ratio)."""
    if abs(float(getattr(model, "downsample_ratio", 0.5)) - 0.5) > 1e-9:
        raise _lib.UllsamError("bridge: kernels are specialised for downsample_ratio 0.5")

    def _extract(self, pixel_values):
        return extract_feature(self.vision_model, self.mlp1, pixel_values, getattr(self, "ps_version", "v2"))

    def _dense(self, features):
        return text_aware_dense_feature(self.mlp2, features, getattr(self, "ps_version", "v2"))

    model.extract_feature = types.MethodType(_extract, model)
    model.text_aware_dense_feature = types.MethodType(_dense, model)
    return model
