"""The nn.Module drop-ins (ullsam_b200.modeling) on a real GPU: same results as the handle-level API, weight
re-upload on change, dtype handling, and post-hoc replacement on a Sam-like container."""
from functools import partial

import pytest
import torch
from torch import nn

from ullsam_b200 import synth
from ullsam_b200.config import config

pytestmark = pytest.mark.gpu


def _build(sd):
    from ullsam_b200.modeling import ImageEncoderViT, MaskDecoder, TwoWayTransformer

    enc = ImageEncoderViT(depth=12, embed_dim=768, img_size=1024, mlp_ratio=4, norm_layer=partial(nn.LayerNorm, eps=1e-6),
                          num_heads=12, patch_size=16, qkv_bias=True, use_rel_pos=True, global_attn_indexes=[2, 5, 8, 11],
                          window_size=14, out_chans=256)
    dec = MaskDecoder(num_multimask_outputs=3, transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048,
                                                                             num_heads=8),
                      transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256)
    enc.load_state_dict({k[len("image_encoder."):]: v for k, v in sd.items() if k.startswith("image_encoder.")})
    dec.load_state_dict({k[len("mask_decoder."):]: v for k, v in sd.items() if k.startswith("mask_decoder.")})
    return enc.cuda().eval(), dec.cuda().eval()


@pytest.fixture(scope="module")
def ctx():
    from ullsam_b200.engine import Engine

    sd = synth.make_state_dict(config("vit_b"), 1234)
    enc, dec = _build(sd)
    eng = Engine(model="vit_b")
    eng.load_state_dict(sd)
    images, _ = synth.make_images(2, 1024, seed=5)
    sparse, pe, dvec = synth.make_prompt_embeddings(2)
    return dict(sd=sd, enc=enc, dec=dec, eng=eng, images=images.cuda(), sparse=sparse.cuda(), pe=pe.cuda(), dvec=dvec.cuda())


def test_modules_match_engine_bitwise(ctx):
    enc, dec, eng = ctx["enc"], ctx["dec"], ctx["eng"]
    with torch.no_grad():
        emb = enc(ctx["images"])
        dense = ctx["dvec"].reshape(1, 256, 1, 1).expand(2, 256, 64, 64)  # PromptEncoder's no_mask_embed broadcast
        masks, iou = dec(emb, ctx["pe"], ctx["sparse"], dense, multimask_output=True)
        m4, s4 = dec.predict_masks(emb[:1], ctx["pe"], ctx["sparse"][:1], dense[:1])
        q, k = dec.transformer(emb[:1] + dense[:1], ctx["pe"], torch.cat([torch.cat([dec.iou_token.weight,
                               dec.mask_tokens.weight])[None], ctx["sparse"][:1]], dim=1))
    emb2 = eng.encoder_forward(ctx["images"])
    masks2, iou2 = eng.decoder_forward(emb2, ctx["pe"], ctx["sparse"], ctx["dvec"][None], True, prompts_per_image=1)
    assert emb.shape == (2, 256, 64, 64) and torch.equal(emb, emb2)
    assert masks.shape == (2, 3, 256, 256) and iou.shape == (2, 3)
    assert torch.equal(masks, masks2) and torch.equal(iou, iou2)
    assert m4.shape == (1, 4, 256, 256) and torch.equal(m4[:, 1:], masks[:1]) and torch.equal(s4[:, 1:], iou[:1])
    assert q.shape == (1, 7, 256) and k.shape == (1, 4096, 256) and torch.isfinite(k).all()


def test_one_image_many_prompts_equals_per_prompt_calls(ctx):
    """Reference calling convention (mask_decoder.py:126): 1 image embedding, P prompt sets."""
    enc, dec = ctx["enc"], ctx["dec"]
    with torch.no_grad():
        emb = enc(ctx["images"][:1])
        dense = ctx["dvec"].reshape(1, 256, 1, 1).expand(2, 256, 64, 64)
        both, _ = dec(emb, ctx["pe"], ctx["sparse"], dense, multimask_output=False)
        one, _ = dec(emb, ctx["pe"], ctx["sparse"][1:2], dense[:1], multimask_output=False)
    assert both.shape == (2, 1, 256, 256) and torch.equal(both[1:2], one)


def test_weight_change_is_picked_up(ctx):
    enc = ctx["enc"]
    with torch.no_grad():
        a = enc(ctx["images"][:1])
        enc.neck[3].bias.add_(1.0)  # in-place: storage stays, version counter moves
        b = enc(ctx["images"][:1])
        enc.neck[3].bias.sub_(1.0)
        c = enc(ctx["images"][:1])
    assert torch.allclose(b, a + 1.0, atol=1e-5) and torch.allclose(c, a, atol=1e-5)  # (x + 1) - 1 != x bitwise


def test_bf16_model_like_the_reference_app(ctx):
    """app.py casts the model to bf16 and feeds bf16 pixels (app.py:328,525)."""
    enc16, _ = _build(ctx["sd"])
    enc16 = enc16.to(torch.bfloat16)
    with torch.no_grad():
        out16 = enc16(ctx["images"][:1].bfloat16())
        out32 = ctx["enc"](ctx["images"][:1])
    assert out16.dtype == torch.bfloat16
    rel = ((out16.float() - out32).norm() / out32.norm()).item()
    print(f"bf16-parameter model vs fp32-parameter model: rel-L2 {rel:.3e}")
    assert rel < 2e-2


def test_accelerate_replaces_modules_on_a_sam_like_container(ctx):
    from ullsam_b200 import modeling as M

    class RefEncoder(nn.Module):  # same attribute tree as the reference's ImageEncoderViT, but a different class
        def __init__(self, src):
            super().__init__()
            self.img_size, self.patch_embed, self.pos_embed = src.img_size, src.patch_embed, src.pos_embed
            self.blocks, self.neck = src.blocks, src.neck

    class RefTransformer(nn.Module):
        def __init__(self, src):
            super().__init__()
            self.depth, self.embedding_dim, self.num_heads, self.mlp_dim = src.depth, src.embedding_dim, src.num_heads, src.mlp_dim
            self.layers, self.final_attn_token_to_image, self.norm_final_attn = src.layers, src.final_attn_token_to_image, src.norm_final_attn

    class RefDecoder(nn.Module):
        def __init__(self, src):
            super().__init__()
            self.transformer_dim, self.num_multimask_outputs = src.transformer_dim, src.num_multimask_outputs
            self.transformer = RefTransformer(src.transformer)
            self.iou_token, self.mask_tokens, self.output_upscaling = src.iou_token, src.mask_tokens, src.output_upscaling
            self.output_hypernetworks_mlps, self.iou_prediction_head = src.output_hypernetworks_mlps, src.iou_prediction_head

    class SamLike(nn.Module):
        def __init__(self, e, d):
            super().__init__()
            self.image_encoder, self.mask_decoder = e, d

    enc, dec = _build(ctx["sd"])
    sam = SamLike(RefEncoder(enc), RefDecoder(dec))
    M.accelerate(sam)
    assert isinstance(sam.image_encoder, M.ImageEncoderViT) and isinstance(sam.mask_decoder, M.MaskDecoder)
    with torch.no_grad():
        emb = sam.image_encoder(ctx["images"][:1])
    assert torch.equal(emb, ctx["eng"].encoder_forward(ctx["images"][:1]))
