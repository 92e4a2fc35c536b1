"""Config 5 of BASELINE.json: the hot path INSIDE the reference's real `InternVLSAMModel`
(modeling/modeling_internvl_sam.py:35-201, 226-270), driven the way train_joint_v2.train_epoch drives it (:959-1079).

TEST INFRASTRUCTURE (oracle side): builds the unmodified reference classes from oracle/_ref (see oracle/make_ref.py)
with seeded synthetic weights — the SAM branch from ullsam_b200.synth (the same state_dict the CUDA engine loads), the
language model a small random-init InternLM2 (InternVL2.5-2B weights are not available offline; the SAM branch does not
depend on the LLM's size, only on the [B, 1024, H] hidden states it hands back).

One shim is needed to construct the reference under transformers >= 4.50: the config round-trip rewrites
`llm_config.rope_scaling` to a dict without the 'type' key that modeling_internlm2.py:318 reads, so it is reset to None
after the config is built.  Nothing in the reference files is edited.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

IMG_CONTEXT_TOKEN_ID = 92546      # modeling_internvl_sam.py:103
N_IMG_TOKENS = 1024               # 64*64 * downsample_ratio**2 (:52)


def llm_config_dict(hidden: int = 256, layers: int = 2, heads: int = 4, kv_heads: int = 2, inter: int = 512,
                    vocab: int = 92553) -> dict:
    return dict(architectures=["InternLM2ForCausalLM"], hidden_size=hidden, num_hidden_layers=layers,
                num_attention_heads=heads, num_key_value_heads=kv_heads, intermediate_size=inter, vocab_size=vocab,
                max_position_embeddings=2048, rms_norm_eps=1e-5, rope_theta=1000000, bias=False, tie_word_embeddings=False,
                attn_implementation="eager", pad_token_id=2, bos_token_id=1, eos_token_id=2)


def build(sd: dict, seed: int = 4321, ps_version: str = "v2", hidden: int = 256, layers: int = 2, model: str = "vit_b"):
    """-> the reference's InternVLSAMModel (eval, fp32, CPU) wrapping the reference's Sam parts loaded from `sd`."""
    from . import make_ref

    modeling, build_sam = make_ref.import_reference()
    from modeling.configuration_internvl_chat import InternVLChatConfig

    torch.manual_seed(seed)
    sam = build_sam.sam_model_registry[model](None)
    missing, unexpected = sam.load_state_dict(sd, strict=False)
    assert not unexpected and all(k in ("pixel_mean", "pixel_std") for k in missing), (missing, unexpected)
    cfg = InternVLChatConfig(llm_config=llm_config_dict(hidden=hidden, layers=layers), downsample_ratio=0.5,
                             ps_version=ps_version, template="internvl2_5", select_layer=-1)
    cfg.llm_config.rope_scaling = None          # the shim described in the module docstring
    cfg.llm_config.attn_implementation = "eager"
    m = modeling.InternVLSAMModel(cfg, vision_model=sam.image_encoder, prompt_encoder=sam.prompt_encoder,
                                  mask_decoder=sam.mask_decoder, use_flash_attn=False)
    # seeded, non-degenerate projector LayerNorms (default init gamma = 1, beta = 0 would hide affine mistakes)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for mlp in (m.mlp1, m.mlp2):
            mlp[0].weight.add_(0.1 * torch.randn(mlp[0].weight.shape, generator=g))
            mlp[0].bias.add_(0.1 * torch.randn(mlp[0].bias.shape, generator=g))
    return m.eval()


def make_inputs(B: int, seed: int = 99, prefix: int = 6, suffix: int = 10) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """input_ids = [text prefix | 1024 x <IMG_CONTEXT> | text suffix] (train_joint_v2.py:300-345 builds the same shape
    with the real tokenizer), attention mask, image_flags."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(5, 30000, (B, prefix + N_IMG_TOKENS + suffix), generator=g)
    ids[:, prefix:prefix + N_IMG_TOKENS] = IMG_CONTEXT_TOKEN_ID
    return ids, torch.ones_like(ids), torch.ones(B, 1, dtype=torch.long)


@torch.no_grad()
def segment(m, pixel_values: torch.Tensor, points: torch.Tensor, point_labels: torch.Tensor, input_ids: torch.Tensor,
            attention_mask: torch.Tensor, image_flags: torch.Tensor, multimask_output: bool,
            reuse_embeddings: bool = False) -> dict:
    """The SAM branch of train_joint_v2.train_epoch (:959-1079) / app.py:497-648 for ONE image with P point prompts:
    model.forward(output_hidden_states=True) -> dense map from the LLM's hidden states -> (second) vision_model pass ->
    prompt_encoder(points, llm_hidden_states) -> mask_decoder.  `reuse_embeddings` takes outputs.image_embeddings
    instead of the reference's second encoder pass (train_joint_v2.py:1015) — numerically the same tensor."""
    out = m(pixel_values=pixel_values, input_ids=input_ids, attention_mask=attention_mask, image_flags=image_flags,
            output_hidden_states=True, return_dict=True)
    dense_map = out.hidden_states                                   # [1, 256, 64, 64]  (:176-190)
    emb = out.image_embeddings if reuse_embeddings else m.vision_model(pixel_values)
    pe = m.prompt_encoder.get_dense_pe().to(emb.device)
    bs = points.shape[0]
    if dense_map.shape[0] != bs:
        dense_map = dense_map.repeat(bs, 1, 1, 1)                  # train_joint_v2.py:1033-1035
    sparse, dense = m.prompt_encoder(points=(points, point_labels), boxes=None, masks=None, llm_hidden_states=dense_map)
    masks, iou = m.mask_decoder(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sparse,
                                dense_prompt_embeddings=dense, multimask_output=multimask_output)
    return dict(masks=masks, iou=iou, emb=emb, dense_map=out.hidden_states, sparse=sparse, dense=dense,
                logits=out.logits)
