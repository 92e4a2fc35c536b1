"""Module-level drop-ins for the reference's ``modeling`` package.

    import modeling                                   # the reference package
    from ullsam_b200.modeling import install
    install(modeling)                                 # before ``import build_sam``: class substitution
    # or, on an already-built model:
    from ullsam_b200.modeling import accelerate
    accelerate(sam)                                   # Sam, or uLLSAM's InternVLSAMModel
"""
from __future__ import annotations

import torch
from torch import nn

from .image_encoder import ImageEncoderViT
from .mask_decoder import MaskDecoder
from .prompt_encoder import PromptEncoder
from .transformer import TwoWayTransformer

__all__ = ["ImageEncoderViT", "MaskDecoder", "PromptEncoder", "TwoWayTransformer", "install", "accelerate", "convert_encoder",
           "convert_decoder", "convert_prompt_encoder"]


def install(reference_modeling_module) -> None:
    """Class substitution: ``build_sam.py`` does ``from modeling import ImageEncoderViT, MaskDecoder, ...`` at import
    time (build_sam.py:11), so replacing the names on the package before that import swaps the path in."""
    for cls in (ImageEncoderViT, MaskDecoder, TwoWayTransformer, PromptEncoder):
        setattr(reference_modeling_module, cls.__name__, cls)


def convert_encoder(ref: nn.Module) -> ImageEncoderViT:
    """Build the CUDA-path encoder from a reference ``ImageEncoderViT`` instance (weights copied via state_dict)."""
    blocks = list(ref.blocks)
    dim = ref.patch_embed.proj.out_channels
    new = ImageEncoderViT(
        img_size=ref.img_size, patch_size=ref.patch_embed.proj.kernel_size[0], in_chans=ref.patch_embed.proj.in_channels,
        embed_dim=dim, depth=len(blocks), num_heads=blocks[0].attn.num_heads,
        mlp_ratio=blocks[0].mlp.lin1.out_features / dim, out_chans=ref.neck[0].out_channels, qkv_bias=True,
        norm_layer=lambda d: nn.LayerNorm(d, eps=blocks[0].norm1.eps), use_abs_pos=ref.pos_embed is not None,
        use_rel_pos=blocks[0].attn.use_rel_pos, window_size=max(b.window_size for b in blocks),
        global_attn_indexes=tuple(i for i, b in enumerate(blocks) if b.window_size == 0))
    new.load_state_dict(ref.state_dict(), strict=True)
    p = next(ref.parameters())
    return new.to(device=p.device, dtype=p.dtype).train(ref.training)


def convert_decoder(ref: nn.Module) -> MaskDecoder:
    """Build the CUDA-path decoder from a reference ``MaskDecoder`` instance."""
    t = ref.transformer
    new = MaskDecoder(
        transformer_dim=ref.transformer_dim,
        transformer=TwoWayTransformer(depth=t.depth, embedding_dim=t.embedding_dim, num_heads=t.num_heads, mlp_dim=t.mlp_dim),
        num_multimask_outputs=ref.num_multimask_outputs, iou_head_depth=ref.iou_prediction_head.num_layers,
        iou_head_hidden_dim=ref.iou_prediction_head.layers[0].out_features)
    new.load_state_dict(ref.state_dict(), strict=True)
    p = next(ref.parameters())
    return new.to(device=p.device, dtype=p.dtype).train(ref.training)


def convert_prompt_encoder(ref: nn.Module) -> PromptEncoder:
    """Build the CUDA-path prompt encoder from a reference ``PromptEncoder`` instance."""
    new = PromptEncoder(embed_dim=ref.embed_dim, image_embedding_size=tuple(ref.image_embedding_size),
                        input_image_size=tuple(ref.input_image_size), mask_in_chans=ref.mask_downscaling[3].out_channels)
    new.load_state_dict(ref.state_dict(), strict=True)
    p = next(ref.parameters())
    return new.to(device=p.device, dtype=p.dtype).train(ref.training)


def accelerate(model: nn.Module) -> nn.Module:
    """Post-hoc module replacement on a built model.  Handles the reference's ``Sam`` (attributes ``image_encoder`` /
    ``mask_decoder``, sam.py:43-45) and uLLSAM's ``InternVLSAMModel`` (``vision_model`` / ``mask_decoder``,
    train_joint_v2.py:1365-1369)."""
    with torch.no_grad():
        for attr in ("image_encoder", "vision_model"):
            if hasattr(model, attr) and not isinstance(getattr(model, attr), ImageEncoderViT):
                setattr(model, attr, convert_encoder(getattr(model, attr)))
        if hasattr(model, "mask_decoder") and not isinstance(model.mask_decoder, MaskDecoder):
            model.mask_decoder = convert_decoder(model.mask_decoder)
        if hasattr(model, "prompt_encoder") and not isinstance(model.prompt_encoder, PromptEncoder):
            model.prompt_encoder = convert_prompt_encoder(model.prompt_encoder)
    return model
