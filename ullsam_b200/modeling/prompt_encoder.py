"""Drop-in for the reference's ``modeling.prompt_encoder.PromptEncoder`` (prompt_encoder.py:16-250), the step
immediately upstream of the mask decoder.

Same constructor, attributes (``embed_dim``, ``input_image_size``, ``image_embedding_size``, ``pe_layer``,
``point_embeddings``, ``not_a_point_embed``, ``llm_scale_factor``, ``llm_bias``, ``mask_downscaling``,
``no_mask_embed``), ``state_dict`` keys, ``forward(points, boxes, masks, llm_hidden_states=None)`` and
``get_dense_pe()``.  The random-Fourier point / box / grid encodings and uLLSAM's LLM-conditioned dense prompt
(``_normalize_llm_hidden_states``) run as CUDA kernels behind the C ABI (``ullsam_op_prompt_*``); the reference's
~20 tiny eager launches and its boolean-mask indexing (a host sync per call) are gone.  ``get_dense_pe`` is cached: it
only depends on a buffer.

The mask-prompt branch of stock SAM (``masks is not None`` -> ``mask_downscaling``, prompt_encoder.py:105-108, 176-177;
uLLSAM itself always passes ``masks=None``) is one fused kernel (``ullsam_op_prompt_mask_embed``), so the class is a
complete drop-in for ``Sam.forward`` (sam.py:106-110) as well.
"""
from __future__ import annotations

from typing import Any, Optional, Tuple, Type

import torch
from torch import nn

from .. import _lib, ops
from .common import LayerNorm2d, _NoEagerForward


class PositionEmbeddingRandom(_NoEagerForward):
    """prompt_encoder.py:206-250: holds the ``positional_encoding_gaussian_matrix`` buffer."""

    def __init__(self, num_pos_feats: int = 64, scale: Optional[float] = None) -> None:
        super().__init__()
        if scale is None or scale <= 0.0:
            scale = 1.0
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn((2, num_pos_feats)))


class PromptEncoder(nn.Module):
    def __init__(self, embed_dim: int, image_embedding_size: Tuple[int, int], input_image_size: Tuple[int, int],
                 mask_in_chans: int, activation: Type[nn.Module] = nn.GELU) -> None:
        super().__init__()
        self.embed_dim = embed_dim
        self.input_image_size = input_image_size
        self.image_embedding_size = image_embedding_size
        self.pe_layer = PositionEmbeddingRandom(embed_dim // 2)
        self.num_point_embeddings: int = 4  # pos/neg point + 2 box corners
        self.point_embeddings = nn.ModuleList([nn.Embedding(1, embed_dim) for _ in range(self.num_point_embeddings)])
        self.not_a_point_embed = nn.Embedding(1, embed_dim)
        self.llm_scale_factor = nn.Parameter(torch.ones(1) * 0.1)
        self.llm_bias = nn.Parameter(torch.zeros(1))
        self.mask_input_size = (4 * image_embedding_size[0], 4 * image_embedding_size[1])
        self.mask_downscaling = nn.Sequential(
            nn.Conv2d(1, mask_in_chans // 4, kernel_size=2, stride=2),
            LayerNorm2d(mask_in_chans // 4),
            activation(),
            nn.Conv2d(mask_in_chans // 4, mask_in_chans, kernel_size=2, stride=2),
            LayerNorm2d(mask_in_chans),
            activation(),
            nn.Conv2d(mask_in_chans, embed_dim, kernel_size=1),
        )
        self.no_mask_embed = nn.Embedding(1, embed_dim)
        self._dense_pe_cache: Optional[Tuple[Any, torch.Tensor]] = None

    # ------------------------------------------------------------------------------------------------------
    def _gauss(self) -> torch.Tensor:
        g = self.pe_layer.positional_encoding_gaussian_matrix
        if not g.is_cuda:
            raise _lib.UllsamError("PromptEncoder (sm_100a) needs its parameters on a CUDA device; there is no CPU path")
        return g

    def _get_device(self) -> torch.device:
        return self.point_embeddings[0].weight.device

    def get_dense_pe(self) -> torch.Tensor:
        """prompt_encoder.py:65-74 -> [1, embed_dim, H, W]; cached per (buffer storage, version)."""
        g = self._gauss()
        if self.image_embedding_size[0] != self.image_embedding_size[1]:
            raise _lib.UllsamError("PromptEncoder (sm_100a): square embedding grids only")
        key = (g.data_ptr(), g._version, g.dtype, g.device)
        if self._dense_pe_cache is None or self._dense_pe_cache[0] != key:
            self._dense_pe_cache = (key, ops.prompt_dense_pe(g, self.image_embedding_size[0]).unsqueeze(0))
        return self._dense_pe_cache[1]

    def _normalize_llm_hidden_states(self, llm_hidden_states: torch.Tensor) -> torch.Tensor:
        """prompt_encoder.py:131-151."""
        return ops.prompt_llm_dense(llm_hidden_states, self.llm_scale_factor.detach(), self.llm_bias.detach())

    def forward(self, points: Optional[Tuple[torch.Tensor, torch.Tensor]], boxes: Optional[torch.Tensor],
                masks: Optional[torch.Tensor], llm_hidden_states: Optional[torch.Tensor] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        """prompt_encoder.py:153-203: -> (sparse [B, N, embed_dim], dense [B, embed_dim, H, W])."""
        # forward only: the same rule as the encoder / decoder shims (ullsam_b200/modeling/_sync.py) — with grad mode on,
        # llm_scale_factor / llm_bias / the point embeddings requiring grad means a backward is expected
        if torch.is_grad_enabled() and (
                (llm_hidden_states is not None and llm_hidden_states.requires_grad)
                or any(p.requires_grad for p in self.parameters())):
            raise _lib.UllsamError("PromptEncoder (sm_100a) is forward-only: grad mode is on and an input or a parameter "
                                   "requires grad; call it under torch.no_grad() or freeze it with .requires_grad_(False)")
        g = self._gauss()
        dev = self._get_device()
        if points is not None:                       # _get_batch_size, prompt_encoder.py:110-125
            bs = points[0].shape[0]
        elif boxes is not None:
            bs = boxes.shape[0]
        elif masks is not None:
            bs = masks.shape[0]
        else:
            bs = 1
        if points is None and boxes is None:
            sparse = torch.empty((bs, 0, self.embed_dim), device=dev)
        else:
            coords, labels = points if points is not None else (None, None)
            pe = torch.cat([e.weight.detach() for e in self.point_embeddings], dim=0)
            sparse = ops.prompt_sparse(g, pe, self.not_a_point_embed.weight.detach(), coords, labels, boxes,
                                       pad=(boxes is None), img_w=self.input_image_size[1], img_h=self.input_image_size[0])
        H, W = self.image_embedding_size
        if masks is not None:                        # takes precedence over the LLM map, as in the reference (:176-177)
            if tuple(masks.shape[-2:]) != tuple(self.mask_input_size):
                raise _lib.UllsamError(f"PromptEncoder (sm_100a): mask prompts must be {self.mask_input_size}, got "
                                       f"{tuple(masks.shape[-2:])}")
            md = self.mask_downscaling
            dense = ops.prompt_mask_embed(masks, [md[0].weight, md[0].bias, md[1].weight, md[1].bias, md[3].weight, md[3].bias,
                                                  md[4].weight, md[4].bias, md[6].weight, md[6].bias])
        elif llm_hidden_states is not None:
            dense = self._normalize_llm_hidden_states(llm_hidden_states).reshape(bs, -1, H, W)
        else:
            dense = self.no_mask_embed.weight.detach().reshape(1, -1, 1, 1).expand(bs, -1, H, W)
        return sparse, dense
