"""Pins the CPU oracle (oracle/sam_oracle.py) against outputs of the reference's own modules
(tests/golden/*.npz, written by oracle/make_golden.py in the build container).  fp32 vs fp32: tolerance 2e-4
relative to the tensor's scale (summation-order differences only)."""
import os

import numpy as np
import pytest
import torch

from oracle import sam_oracle as O
from ullsam_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _close(got, ref, what, tol=2e-4):
    got = got.detach().numpy() if isinstance(got, torch.Tensor) else got
    scale = max(float(np.abs(ref).max()), 1e-6)
    err = float(np.abs(got - ref).max())
    assert err <= tol * scale, f"{what}: max|err|={err:.3e} scale={scale:.3e}"


@pytest.fixture(scope="module", params=["tiny", "vit_b"])
def case(request):
    name = request.param
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    cfg = O.config(name)
    seed, B = int(g["seed"]), int(g["B"])
    sd = synth.make_state_dict(cfg, seed)
    images, boxes = synth.make_images(B, cfg.img_size, seed=seed + 1)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        emb = O.encoder_forward(sd, images, cfg)
    return dict(name=name, g=g, cfg=cfg, sd=sd, seed=seed, B=B, images=images, boxes=boxes, emb=emb)


def test_encoder_matches_reference(case):
    g, es = case["g"], int(case["g"]["emb_stride"])
    _close(case["emb"][:, :, ::es, ::es], g["emb"], "image embedding")
    _close(case["emb"].pow(2).sum(dim=(1, 2, 3)).sqrt(), g["emb_l2"], "embedding L2 norm")


def test_dense_pe_matches_reference(case):
    pe = O.dense_pe(case["sd"], case["cfg"])
    _close(pe[0, :, ::8, ::8], case["g"]["image_pe"], "dense positional encoding", tol=1e-5)


@pytest.mark.parametrize("multimask", [False, True])
def test_box_prompt_decoder_matches_reference(case, multimask):
    g, cfg, sd = case["g"], case["cfg"], case["sd"]
    ms = int(g["mask_stride"])
    tag = "multi" if multimask else "single"
    pe = O.dense_pe(sd, cfg)
    with torch.no_grad():
        sparse = O.embed_boxes(sd, case["boxes"], cfg)
        _close(sparse[-1:], g["box_sparse"], "box corner embeddings", tol=1e-5)
        masks, ious = [], []
        for i in range(case["B"]):
            m, s = O.decoder_forward(sd, case["emb"][i:i + 1], pe, sparse[i:i + 1], O.dense_no_mask(sd, 1, cfg),
                                     multimask, cfg)
            masks.append(m)
            ious.append(s)
    masks, ious = torch.cat(masks), torch.cat(ious)
    assert masks.shape[1] == (3 if multimask else 1)
    _close(masks[:, :, ::ms, ::ms], g[f"box_masks_{tag}"], f"mask logits ({tag})")
    _close(ious, g[f"box_iou_{tag}"], f"iou predictions ({tag})")
    _close((masks > 0).float().mean(dim=(2, 3)), g[f"box_pos_frac_{tag}"], "foreground fraction", tol=2e-3)


def test_llm_conditioned_decoder_matches_reference(case):
    """uLLSAM calling convention: one image, P point-prompt instances, LLM-derived dense prompt."""
    g, cfg, sd, seed = case["g"], case["cfg"], case["sd"], case["seed"]
    ms = int(g["mask_stride"])
    P = g["llm_masks"].shape[0]
    coords, labels = synth.make_points(P, cfg.img_size, seed=seed + 2)
    llm = synth.make_llm_map(1, seed=seed + 3, C=cfg.transformer_dim, g=cfg.grid).repeat(P, 1, 1, 1)
    with torch.no_grad():
        sparse = O.embed_points(sd, coords, labels, cfg, pad=True)
        dense = O.dense_from_llm(sd, llm)
        _close(sparse, g["llm_sparse"], "point embeddings", tol=1e-5)
        _close(dense[:, :, ::8, ::8], g["llm_dense"], "LLM dense prompt", tol=1e-5)
        pe = O.dense_pe(sd, cfg)
        stages = {}
        masks, iou = O.predict_masks(sd, case["emb"][0:1], pe, sparse, dense, cfg, stages=stages)
    _close(masks[:, 0:1, ::ms, ::ms], g["llm_masks"], "LLM-conditioned mask logits")
    _close(iou[:, 0:1], g["llm_iou"], "LLM-conditioned iou")
    _close(stages["queries"], g["llm_queries"], "transformer queries")
    _close(stages["keys"][:, ::64], g["llm_keys"], "transformer keys")


def test_window_partition_roundtrip_and_padding():
    x = torch.arange(2 * 20 * 20 * 3, dtype=torch.float32).reshape(2, 20, 20, 3) + 1
    win, pad_hw = O.window_partition(x, 14)
    assert win.shape == (8, 14, 14, 3) and pad_hw == (28, 28)
    assert (win[1, :, 6:] == 0).all()  # right-edge window: columns 20..27 are padding
    assert torch.equal(O.window_unpartition(win, 14, pad_hw, (20, 20)), x)


def test_rel_pos_table_indexing():
    rp = torch.arange(27, dtype=torch.float32)[:, None].expand(27, 4)
    R = O.rel_pos_table(14, rp)
    assert R[0, 0, 0] == 13 and R[13, 0, 0] == 26 and R[0, 13, 0] == 0


def test_bridge_restatement_vs_reference_golden():
    """oracle pixel_shuffle / mllm_extract_feature / text_aware_dense_feature against outputs of the reference's own
    InternVLSAMModel methods (oracle/make_golden_bridge.py)."""
    import os
    import numpy as np
    import torch
    from oracle import sam_oracle as O
    from oracle.make_golden_bridge import make_mlp

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bridge.npz"))
    H = int(g["H"])
    emb = torch.randn(1, 256, 64, 64, generator=torch.Generator().manual_seed(11))
    hid = torch.randn(1, 1024, H, generator=torch.Generator().manual_seed(12))
    p1 = {k: v.detach() for k, v in make_mlp(1024, H, H, 21).state_dict().items()}
    p2 = {k: v.detach() for k, v in make_mlp(H, 1024, 1024, 22).state_dict().items()}
    with torch.no_grad():
        for v in ("v2", "v1"):
            assert np.array_equal(O.pixel_shuffle(emb.permute(0, 2, 3, 1), 0.5, v)[:, ::4, ::4, ::8].numpy(), g[f"shuffle_{v}"])
            tok = O.mllm_extract_feature(emb, p1, v)
            assert np.abs(tok[:, ::16].numpy() - g[f"tokens_{v}"]).max() < 1e-5
            dense = O.text_aware_dense_feature(hid, p2, v)
            assert np.abs(dense[:, :, ::4, ::4].numpy() - g[f"dense_{v}"]).max() < 1e-5


def test_postprocess_restatement_vs_reference_golden():
    """oracle postprocess_masks / calc_iou against the reference's Sam.postprocess_masks and CalcIoU
    (oracle/make_golden_post.py)."""
    import os
    import numpy as np
    import torch
    from oracle import sam_oracle as O
    from oracle.make_golden_post import CASES, low_res

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "post.npz"))
    for i, (inp, orig) in enumerate(CASES):
        m = O.postprocess_masks(low_res(30 + i), inp, orig)
        assert np.array_equal(m[:, :, ::13, ::13].numpy(), g[f"post_{i}"])
    logits = torch.nn.functional.interpolate(low_res(40), (512, 512), mode="bilinear", align_corners=False)
    target = (torch.nn.functional.interpolate(low_res(41), (512, 512), mode="bilinear", align_corners=False) > 0.5).float()
    assert np.allclose(O.calc_iou(logits, target).numpy(), g["iou"], atol=1e-7)


def test_mask_prompt_branch_vs_reference_golden():
    """oracle embed_masks (stock SAM's mask_downscaling, prompt_encoder.py:53-61, 105-108) and the decoder fed with it,
    against the reference's own PromptEncoder / MaskDecoder (oracle/make_golden_prompt_mask.py)."""
    import os
    import numpy as np
    import torch
    from oracle import sam_oracle as O
    from oracle.make_golden_prompt_mask import make_mask_inputs
    from ullsam_b200 import synth

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "prompt_mask.npz"))
    seed, P = int(g["seed"]), int(g["P"])
    cfg = O.config("vit_b")
    sd = synth.make_state_dict(cfg, seed)
    masks_in = make_mask_inputs(P, seed + 21)
    _, boxes = synth.make_images(P, 1024, seed=seed + 1)
    with torch.no_grad():
        dense = O.embed_masks(sd, masks_in)
        assert dense.shape == (P, 256, 64, 64)
        assert np.abs(dense[:, :, ::4, ::4].numpy() - g["dense"]).max() < 1e-5
        sparse = O.embed_boxes(sd, boxes, cfg)
        assert np.abs(sparse.numpy() - g["sparse"]).max() < 1e-5
        emb = torch.randn(1, 256, 64, 64, generator=torch.Generator().manual_seed(seed + 22)) * 0.5
        m, s = O.decoder_forward(sd, emb, O.dense_pe(sd, cfg), sparse, dense, True, cfg)
        assert np.abs(m[:, :, ::4, ::4].numpy() - g["masks"]).max() < 2e-3 and np.abs(s.numpy() - g["iou"]).max() < 2e-4
