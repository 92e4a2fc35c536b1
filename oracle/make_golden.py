"""Generate tests/golden/*.npz by running the REFERENCE's own modules (imported from /root/reference, which only
exists in the build container) on the seeded synthetic weights/inputs of ullsam_b200/synth.py.

    python -m oracle.make_golden            # writes tests/golden/{tiny,vit_b}.npz

The fixtures are what pins the CPU oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.  Only
sub-sampled outputs are stored so the files stay small; the sampling grid is recorded in the file.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

REF = os.environ.get("ULLSAM_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def build_reference(cfg):
    """Instantiate the reference modules with cfg's hyper-parameters (as build_sam._build_sam does)."""
    sys.path.insert(0, REF)
    from functools import partial

    from modeling import ImageEncoderViT, MaskDecoder, PromptEncoder, Sam, TwoWayTransformer  # noqa: E402

    g = cfg.grid
    sam = Sam(
        image_encoder=ImageEncoderViT(
            depth=cfg.depth, embed_dim=cfg.embed_dim, img_size=cfg.img_size, mlp_ratio=cfg.mlp_ratio,
            norm_layer=partial(torch.nn.LayerNorm, eps=cfg.enc_ln_eps), num_heads=cfg.num_heads,
            patch_size=cfg.patch_size, qkv_bias=True, use_rel_pos=True,
            global_attn_indexes=list(cfg.global_attn_indexes), window_size=cfg.window_size, out_chans=cfg.out_chans),
        prompt_encoder=PromptEncoder(embed_dim=cfg.transformer_dim, image_embedding_size=(g, g),
                                     input_image_size=(cfg.img_size, cfg.img_size), mask_in_chans=cfg.mask_in_chans),
        mask_decoder=MaskDecoder(
            num_multimask_outputs=cfg.num_multimask_outputs,
            transformer=TwoWayTransformer(depth=cfg.dec_depth, embedding_dim=cfg.transformer_dim,
                                          mlp_dim=cfg.dec_mlp_dim, num_heads=cfg.dec_heads),
            transformer_dim=cfg.transformer_dim, iou_head_depth=cfg.iou_head_depth,
            iou_head_hidden_dim=cfg.iou_head_hidden_dim),
    )
    return sam.eval()


@torch.no_grad()
def run(name: str, B: int, seed: int, emb_stride: int, mask_stride: int):
    from oracle import sam_oracle as O
    from ullsam_b200 import synth

    cfg = O.config(name)
    sd = synth.make_state_dict(cfg, seed)
    sam = build_reference(cfg)
    missing, unexpected = sam.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k in ("pixel_mean", "pixel_std") for k in missing), missing

    images, boxes = synth.make_images(B, cfg.img_size, seed=seed + 1)
    t0 = time.time()
    emb = sam.image_encoder(images)
    t_enc = time.time() - t0
    pe = sam.prompt_encoder.get_dense_pe()

    out = {"emb_stride": emb_stride, "mask_stride": mask_stride, "seed": seed, "B": B,
           "emb": emb[:, :, ::emb_stride, ::emb_stride].numpy(),
           "emb_l2": emb.pow(2).sum(dim=(1, 2, 3)).sqrt().numpy(), "image_pe": pe[0, :, ::8, ::8].numpy()}

    # (a) stock SAM box prompt, dense = no_mask_embed; single and multimask (Sam.forward loop, sam.py:101-117)
    for mm in (False, True):
        ms, ss = [], []
        for i in range(B):
            sp, de = sam.prompt_encoder(points=None, boxes=boxes[i:i + 1], masks=None)
            m, s = sam.mask_decoder(image_embeddings=emb[i:i + 1], image_pe=pe, sparse_prompt_embeddings=sp,
                                    dense_prompt_embeddings=de, multimask_output=mm)
            ms.append(m)
            ss.append(s)
        tag = "multi" if mm else "single"
        m = torch.cat(ms)
        out[f"box_masks_{tag}"] = m[:, :, ::mask_stride, ::mask_stride].numpy()
        out[f"box_pos_frac_{tag}"] = (m > 0).float().mean(dim=(2, 3)).numpy()
        out[f"box_iou_{tag}"] = torch.cat(ss).numpy()
        if not mm:
            out["box_sparse"] = sp.numpy()

    # (b) uLLSAM: P point-prompt instances on ONE image with the LLM-conditioned dense prompt
    #     (train_joint_v2.py:1033-1069: emb[1], sparse [P,5,256], dense [P,256,g,g], multimask False)
    P = 3
    coords, labels = synth.make_points(P, cfg.img_size, seed=seed + 2)
    llm = synth.make_llm_map(1, seed=seed + 3, C=cfg.transformer_dim, g=cfg.grid).repeat(P, 1, 1, 1)
    sp, de = sam.prompt_encoder(points=(coords, labels), boxes=None, masks=None, llm_hidden_states=llm)
    m, s = sam.mask_decoder(image_embeddings=emb[0:1], image_pe=pe, sparse_prompt_embeddings=sp,
                            dense_prompt_embeddings=de, multimask_output=False)
    out["llm_masks"] = m[:, :, ::mask_stride, ::mask_stride].numpy()
    out["llm_iou"] = s.numpy()
    out["llm_sparse"] = sp.numpy()
    out["llm_dense"] = de[:, :, ::8, ::8].numpy()
    # transformer alone (TwoWayTransformer.forward) on the same inputs
    tokens = torch.cat([torch.cat([sam.mask_decoder.iou_token.weight, sam.mask_decoder.mask_tokens.weight])[None]
                        .expand(P, -1, -1), sp], dim=1)
    src = emb[0:1].repeat_interleave(P, 0) + de
    hs, keys = sam.mask_decoder.transformer(src, pe.repeat_interleave(P, 0), tokens)
    out["llm_queries"] = hs.numpy()
    out["llm_keys"] = keys[:, ::64, :].numpy()
    out["ref_encoder_seconds"] = t_enc
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB, reference encoder {t_enc:.2f}s for B={B}")


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    torch.set_num_threads(os.cpu_count() or 1)
    run("tiny", B=2, seed=1234, emb_stride=1, mask_stride=1)
    run("vit_b", B=1, seed=1234, emb_stride=4, mask_stride=4)
