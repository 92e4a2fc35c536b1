"""Drop-in for the reference's ``modeling.transformer.TwoWayTransformer`` (transformer.py:16-108).

Same constructor, attribute tree and ``state_dict`` keys.  ``forward`` runs through ``ullsam_transformer_forward``;
inside a ``MaskDecoder`` the parent's fused call is used instead and this module only holds the parameters.
"""
from __future__ import annotations

from typing import Tuple, Type

import torch
from torch import Tensor, nn

from .. import _lib
from ..engine import make_config
from ._sync import EngineBound
from .common import MLPBlock, _NoEagerForward


class Attention(_NoEagerForward):
    """transformer.py:187-218: q/k/v/out projections with an optional internal down-scaling."""

    def __init__(self, embedding_dim: int, num_heads: int, downsample_rate: int = 1) -> None:
        super().__init__()
        self.embedding_dim = embedding_dim
        self.internal_dim = embedding_dim // downsample_rate
        self.num_heads = num_heads
        assert self.internal_dim % num_heads == 0, "num_heads must divide embedding_dim."
        self.q_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.k_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.v_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.out_proj = nn.Linear(self.internal_dim, embedding_dim)


class TwoWayAttentionBlock(_NoEagerForward):
    """transformer.py:111-151."""

    def __init__(self, embedding_dim: int, num_heads: int, mlp_dim: int = 2048, activation: Type[nn.Module] = nn.ReLU,
                 attention_downsample_rate: int = 2, skip_first_layer_pe: bool = False) -> None:
        super().__init__()
        self.self_attn = Attention(embedding_dim, num_heads)
        self.norm1 = nn.LayerNorm(embedding_dim)
        self.cross_attn_token_to_image = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm2 = nn.LayerNorm(embedding_dim)
        self.mlp = MLPBlock(embedding_dim, mlp_dim, activation)
        self.norm3 = nn.LayerNorm(embedding_dim)
        self.norm4 = nn.LayerNorm(embedding_dim)
        self.cross_attn_image_to_token = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.skip_first_layer_pe = skip_first_layer_pe


class TwoWayTransformer(EngineBound, nn.Module):
    def __init__(self, depth: int, embedding_dim: int, num_heads: int, mlp_dim: int, activation: Type[nn.Module] = nn.ReLU,
                 attention_downsample_rate: int = 2) -> None:
        super().__init__()
        if embedding_dim != 256 or num_heads != 8 or attention_downsample_rate != 2 or activation is not nn.ReLU \
                or mlp_dim > 2048 or mlp_dim % 4:
            raise _lib.UllsamError("TwoWayTransformer (sm_100a): kernels are specialised for SAM's decoder "
                                   "(dim 256, 8 heads, downsample 2, ReLU, mlp_dim <= 2048)")
        self.depth = depth
        self.embedding_dim = embedding_dim
        self.num_heads = num_heads
        self.mlp_dim = mlp_dim
        self.layers = nn.ModuleList()
        for i in range(depth):
            self.layers.append(TwoWayAttentionBlock(embedding_dim=embedding_dim, num_heads=num_heads, mlp_dim=mlp_dim,
                                                    activation=activation,
                                                    attention_downsample_rate=attention_downsample_rate,
                                                    skip_first_layer_pe=(i == 0)))
        self.final_attn_token_to_image = Attention(embedding_dim, num_heads, downsample_rate=attention_downsample_rate)
        self.norm_final_attn = nn.LayerNorm(embedding_dim)

    def _engine_config(self):
        return make_config(depth=0, dec_depth=self.depth, dec_mlp_dim=self.mlp_dim)

    def _key_prefix(self) -> str:
        return "mask_decoder.transformer."

    def forward(self, image_embedding: Tensor, image_pe: Tensor, point_embedding: Tensor) -> Tuple[Tensor, Tensor]:
        """(B x 256 x 64 x 64, same, B x N_points x 256) -> (queries B x N x 256, keys B x 4096 x 256)."""
        self._check_no_grad(image_embedding, image_pe, point_embedding)
        eng = self._get_engine(image_embedding.device)
        q, k = eng.transformer_forward(image_embedding, image_pe, point_embedding)
        dt = image_embedding.dtype
        return (q.to(dt), k.to(dt)) if dt != torch.float32 else (q, k)
