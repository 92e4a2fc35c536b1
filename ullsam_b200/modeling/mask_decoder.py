"""Drop-in for the reference's ``modeling.mask_decoder.MaskDecoder`` (mask_decoder.py:16-149).

Same constructor (keyword-only), attributes (``transformer``, ``num_mask_tokens``, ``iou_token``, ``mask_tokens``,
``output_upscaling``, ``output_hypernetworks_mlps``, ``iou_prediction_head``), ``state_dict`` keys, ``forward`` and
``predict_masks`` signatures.  The whole forward (token assembly, TwoWayTransformer, ConvTranspose upscaling,
hypernetwork MLPs, mask contraction, IoU head, output slicing) is one C-ABI call: ``ullsam_decoder_forward``.
"""
from __future__ import annotations

from typing import Tuple, Type

import torch
from torch import nn

from .. import _lib
from ..engine import make_config
from ._sync import EngineBound
from .common import LayerNorm2d, _NoEagerForward


class MLP(_NoEagerForward):
    """mask_decoder.py:154-176 (Linear/ReLU stack)."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_layers: int, sigmoid_output: bool = False):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))
        self.sigmoid_output = sigmoid_output


class MaskDecoder(EngineBound, nn.Module):
    def __init__(self, *, transformer_dim: int, transformer: nn.Module, num_multimask_outputs: int = 3,
                 activation: Type[nn.Module] = nn.GELU, iou_head_depth: int = 3, iou_head_hidden_dim: int = 256) -> None:
        super().__init__()
        if transformer_dim != 256 or num_multimask_outputs != 3 or activation is not nn.GELU or iou_head_hidden_dim != 256:
            raise _lib.UllsamError("MaskDecoder (sm_100a): kernels are specialised for SAM's decoder "
                                   "(transformer_dim 256, 3 multimask outputs, GELU, iou hidden 256)")
        self.transformer_dim = transformer_dim
        self.transformer = transformer
        self.num_multimask_outputs = num_multimask_outputs
        self.iou_token = nn.Embedding(1, transformer_dim)
        self.num_mask_tokens = num_multimask_outputs + 1
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, transformer_dim)
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(transformer_dim, transformer_dim // 4, kernel_size=2, stride=2),
            LayerNorm2d(transformer_dim // 4),
            activation(),
            nn.ConvTranspose2d(transformer_dim // 4, transformer_dim // 8, kernel_size=2, stride=2),
            activation(),
        )
        self.output_hypernetworks_mlps = nn.ModuleList(
            [MLP(transformer_dim, transformer_dim, transformer_dim // 8, 3) for _ in range(self.num_mask_tokens)])
        self.iou_prediction_head = MLP(transformer_dim, iou_head_hidden_dim, self.num_mask_tokens, iou_head_depth)
        self._iou_head_depth = iou_head_depth

    def _engine_config(self):
        t = self.transformer
        return make_config(depth=0, dec_depth=getattr(t, "depth", 2), dec_mlp_dim=getattr(t, "mlp_dim", 2048),
                           iou_head_depth=self._iou_head_depth)

    def _key_prefix(self) -> str:
        return "mask_decoder."

    def _run(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, multimask):
        self._check_no_grad(image_embeddings, sparse_prompt_embeddings, dense_prompt_embeddings)
        eng = self._get_engine(image_embeddings.device)
        n_img, R = image_embeddings.shape[0], sparse_prompt_embeddings.shape[0]
        if n_img == 1:
            P = R  # the reference's own convention: one image, R prompts (mask_decoder.py:126)
        elif R == n_img:
            P = 1  # concatenation of per-image calls (Sam.forward's loop, sam.py:101-117)
        else:
            raise _lib.UllsamError(f"MaskDecoder: {n_img} image embeddings with {R} prompt sets is not a valid batch "
                                   "(use 1 image x P prompts, or B images x 1 prompt each)")
        masks, iou = eng.decoder_forward(image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings,
                                         multimask, prompts_per_image=P)
        dt = image_embeddings.dtype
        if dt in (torch.bfloat16, torch.float16):
            masks, iou = masks.to(dt), iou.to(dt)
        return masks, iou

    def forward(self, image_embeddings: torch.Tensor, image_pe: torch.Tensor, sparse_prompt_embeddings: torch.Tensor,
                dense_prompt_embeddings: torch.Tensor, multimask_output: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._run(image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings,
                         bool(multimask_output))

    def predict_masks(self, image_embeddings: torch.Tensor, image_pe: torch.Tensor,
                      sparse_prompt_embeddings: torch.Tensor, dense_prompt_embeddings: torch.Tensor
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._run(image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, None)
