"""Drop-in for the reference's ``modeling.image_encoder.ImageEncoderViT`` (image_encoder.py:17-116).

Same constructor signature, attribute names (``img_size``, ``patch_embed``, ``pos_embed``, ``blocks``, ``neck``) and
``state_dict`` keys; ``forward`` runs the whole encoder as hand-written sm_100a kernels through the C ABI
(``ullsam_encoder_forward``).  Forward only, CUDA only — anything else raises.
"""
from __future__ import annotations

from typing import Optional, Tuple, Type

import torch
import torch.nn as nn

from .. import _lib
from ..engine import make_config
from ._sync import EngineBound
from .common import LayerNorm2d, MLPBlock, _NoEagerForward


class PatchEmbed(_NoEagerForward):
    """image_encoder.py:364-395."""

    def __init__(self, kernel_size=(16, 16), stride=(16, 16), padding=(0, 0), in_chans: int = 3, embed_dim: int = 768):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride, padding=padding)


class Attention(_NoEagerForward):
    """image_encoder.py:185-222."""

    def __init__(self, dim: int, num_heads: int = 8, qkv_bias: bool = True, use_rel_pos: bool = False,
                 rel_pos_zero_init: bool = True, input_size: Optional[Tuple[int, int]] = None) -> None:
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.use_rel_pos = use_rel_pos
        if use_rel_pos:
            assert input_size is not None, "Input size must be provided if using relative positional encoding."
            self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size[0] - 1, head_dim))
            self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size[1] - 1, head_dim))


class Block(_NoEagerForward):
    """image_encoder.py:119-164."""

    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0, qkv_bias: bool = True,
                 norm_layer: Type[nn.Module] = nn.LayerNorm, act_layer: Type[nn.Module] = nn.GELU, use_rel_pos: bool = False,
                 rel_pos_zero_init: bool = True, window_size: int = 0, input_size: Optional[Tuple[int, int]] = None) -> None:
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, use_rel_pos=use_rel_pos,
                              rel_pos_zero_init=rel_pos_zero_init,
                              input_size=input_size if window_size == 0 else (window_size, window_size))
        self.norm2 = norm_layer(dim)
        self.mlp = MLPBlock(embedding_dim=dim, mlp_dim=int(dim * mlp_ratio), act=act_layer)
        self.window_size = window_size


class ImageEncoderViT(EngineBound, nn.Module):
    def __init__(self, img_size: int = 1024, patch_size: int = 16, in_chans: int = 3, embed_dim: int = 768, depth: int = 12,
                 num_heads: int = 12, mlp_ratio: float = 4.0, out_chans: int = 256, qkv_bias: bool = True,
                 norm_layer: Type[nn.Module] = nn.LayerNorm, act_layer: Type[nn.Module] = nn.GELU, use_abs_pos: bool = True,
                 use_rel_pos: bool = False, rel_pos_zero_init: bool = True, window_size: int = 0,
                 global_attn_indexes: Tuple[int, ...] = ()) -> None:
        super().__init__()
        # The kernels are specialised for the configuration build_sam.py fixes (build_sam.py:62-81).
        unsupported = []
        if img_size != 1024 or patch_size != 16 or in_chans != 3: unsupported.append("img_size/patch_size/in_chans != 1024/16/3")
        if embed_dim != 64 * num_heads or embed_dim not in (768, 1024):
            unsupported.append("embed_dim/num_heads (64-wide heads with embed_dim 768 = vit_b or 1024 = vit_l; vit_h's "
                               "80-wide heads are not implemented)")
        if out_chans != 256: unsupported.append("out_chans != 256")
        if not (qkv_bias and use_abs_pos and use_rel_pos): unsupported.append("qkv_bias/use_abs_pos/use_rel_pos must be True")
        if window_size != 14: unsupported.append("window_size != 14")
        if act_layer is not nn.GELU: unsupported.append("act_layer != nn.GELU")
        if unsupported:
            raise _lib.UllsamError("ImageEncoderViT (sm_100a): unsupported configuration: " + "; ".join(unsupported))
        self.img_size = img_size
        self.patch_embed = PatchEmbed(kernel_size=(patch_size, patch_size), stride=(patch_size, patch_size),
                                      in_chans=in_chans, embed_dim=embed_dim)
        self.pos_embed: Optional[nn.Parameter] = nn.Parameter(
            torch.zeros(1, img_size // patch_size, img_size // patch_size, embed_dim))
        self.blocks = nn.ModuleList()
        g = img_size // patch_size
        for i in range(depth):
            self.blocks.append(Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                                     norm_layer=norm_layer, act_layer=act_layer, use_rel_pos=use_rel_pos,
                                     rel_pos_zero_init=rel_pos_zero_init,
                                     window_size=window_size if i not in global_attn_indexes else 0, input_size=(g, g)))
        self.neck = nn.Sequential(nn.Conv2d(embed_dim, out_chans, kernel_size=1, bias=False), LayerNorm2d(out_chans),
                                  nn.Conv2d(out_chans, out_chans, kernel_size=3, padding=1, bias=False),
                                  LayerNorm2d(out_chans))
        eps = getattr(self.blocks[0].norm1, "eps", 1e-5) if depth else 1e-6
        self._cfg_kwargs = dict(embed_dim=embed_dim, depth=depth, num_heads=num_heads,
                                global_attn_indexes=tuple(global_attn_indexes), mlp_ratio=mlp_ratio, enc_ln_eps=float(eps))

    def _engine_config(self):
        return make_config(**self._cfg_kwargs)

    def _key_prefix(self) -> str:
        return "image_encoder."

    def _sync_weights(self, eng) -> None:
        # only the encoder's tensors live in this module
        super()._sync_weights(eng)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[B,3,1024,1024] -> [B,256,64,64]; output dtype follows the parameters' dtype (bf16 model => bf16 out)."""
        self._check_no_grad(x)
        eng = self._get_engine(x.device)
        out_dtype = torch.bfloat16 if self.pos_embed.dtype == torch.bfloat16 else torch.float32
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        return eng.encoder_forward(x, out_dtype=out_dtype)
