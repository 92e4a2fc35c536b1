"""Config 5 of BASELINE.json on the GPU: the hot path swapped into the reference's REAL `InternVLSAMModel`
(modeling_internvl_sam.py:35-270) — InternLM2 in PyTorch, the SAM branch (vision_model, prompt_encoder, mask_decoder and
the two projector MLPs) on the sm_100a kernels — with multimask output, driven like train_joint_v2.train_epoch (:959-1079).

Two layers of evidence:
  * against the committed golden made from the reference model on CPU (oracle/make_golden_mllm.py) — needs nothing but
    the fixture: Engine encoder + PromptEncoder shim (LLM-conditioned dense prompt) + decoder, multimask and single;
  * live, when oracle/_ref is staged (it travels to the GPU box with the snapshot): the same reference class is built twice
    from the same seeds, one copy stays stock on the CPU, the other gets `accelerate()` + `patch_internvl_sam()` and runs
    on cuda:0 WITHOUT the second encoder pass (outputs.image_embeddings is reused).
Bars (north_star): embedding rel-L2 <= 1e-2, binarised-mask IoU >= 0.99.
"""
import os
import warnings

import numpy as np
import pytest
import torch

from ullsam_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _iou(a, b):
    a, b = a > 0, b > 0
    return (a & b).sum().item() / max((a | b).sum().item(), 1)


def test_sam_branch_vs_real_internvl_sam_model_golden():
    from oracle import sam_oracle as O
    from ullsam_b200.engine import Engine
    from ullsam_b200.modeling import PromptEncoder

    g = np.load(os.path.join(GOLD, "mllm.npz"))
    seed, P = int(g["seed"]), int(g["P"])
    cfg = O.config("vit_b")
    sd = synth.make_state_dict(cfg, seed)
    images, _ = synth.make_images(1, 1024, seed=seed + 11)
    pts, lab = synth.make_points(P, 1024, seed=seed + 12)
    dense_map = torch.from_numpy(g["dense_map_f16"]).float().cuda()
    eng = Engine(model="vit_b")
    eng.load_state_dict(sd)
    pe_mod = PromptEncoder(256, (64, 64), (1024, 1024), 16)
    pe_mod.load_state_dict({k[len("prompt_encoder."):]: v for k, v in sd.items() if k.startswith("prompt_encoder.")})
    pe_mod = pe_mod.cuda().eval()
    with torch.no_grad():
        emb = eng.encoder_forward(images.cuda())
        ref = torch.from_numpy(g["emb"]).double()
        rel = ((emb[:, :, ::4, ::4].cpu().double() - ref).norm() / ref.norm()).item()
        sparse, dense = pe_mod(points=(pts.cuda(), lab.cuda()), boxes=None, masks=None, llm_hidden_states=dense_map.repeat(P, 1, 1, 1))
        print(f"config 5 vs reference golden: embedding rel-L2 {rel:.3e}")
        assert rel <= 1e-2
        inter = union = 0
        for tag, mm in (("multi", True), ("single", False)):
            m, s = eng.decoder_forward(emb, pe_mod.get_dense_pe(), sparse, dense, mm)
            assert m.shape == (P, 3 if mm else 1, 256, 256)
            ref_bits = torch.from_numpy(np.unpackbits(g[f"bits_{tag}"])[:m.numel()].reshape(tuple(m.shape)).astype(bool))
            mb = m.cpu() > 0
            i_, u_ = (mb & ref_bits).sum().item(), (mb | ref_bits).sum().item()   # full-resolution binarised masks
            inter, union = inter + i_, union + u_
            lerr = (m[:, :, ::4, ::4].cpu() - torch.from_numpy(g[f"masks_{tag}"])).abs().max().item()
            err = (s.cpu() - torch.from_numpy(g[f"iou_{tag}"])).abs().max().item()
            print(f"  {tag}: binarised-mask IoU {i_ / max(u_, 1):.4f}, mask logits max|err| {lerr:.2e} "
                  f"(|logit| max {np.abs(g[f'masks_{tag}']).max():.2f}), iou-head err {err:.2e}")
            # per-slice floor.  This fixture's masks come from random-init weights: the logits are ~N(0, 0.07) with 1.4-2.7 %
            # of the pixels inside +-2e-3 of the threshold, so the slice IoU is set by the ENCODER's bf16 error alone — the
            # exact fp32 decoder (oracle) fed with this library's embedding (rel-L2 6.7e-3, bar 1e-2) scores 0.9958 (multi)
            # and 0.9905 (single) on the same fixture; this library's decoder fed with the exact embedding 0.9982 / 0.9953
            # (tools/exp_config5_error.py).
            assert i_ / max(u_, 1) >= 0.985 and err < 2e-2
        iou = inter / max(union, 1)
        print(f"  all {4 * P} masks of the fixture (predict_masks' four outputs per instance): binarised-mask IoU {iou:.4f} (bar 0.99)")
        assert iou >= 0.99
    eng.close()


def test_real_internvl_sam_model_accelerated_vs_stock():
    from oracle import make_ref

    if not make_ref.available():
        pytest.skip("oracle/_ref not staged (python -m oracle.make_ref in the build container; it ships with the snapshot)")
    warnings.filterwarnings("ignore")
    from oracle import mllm_ref, sam_oracle as O
    from ullsam_b200.bridge import patch_internvl_sam
    from ullsam_b200.modeling import ImageEncoderViT, MaskDecoder, PromptEncoder, accelerate

    cfg = O.config("vit_b")
    sd = synth.make_state_dict(cfg, 1234)
    images, _ = synth.make_images(1, 1024, seed=1245)
    P = 3
    pts, lab = synth.make_points(P, 1024, seed=1246)
    ids, am, fl = mllm_ref.make_inputs(1, seed=1247)

    torch.set_num_threads(os.cpu_count() or 1)
    stock = mllm_ref.build(sd, seed=4321)                       # the reference, untouched, fp32 on the CPU
    ref = mllm_ref.segment(stock, images, pts, lab, ids, am, fl, multimask_output=True)

    fast = mllm_ref.build(sd, seed=4321)                        # same seeds -> same weights
    fast = patch_internvl_sam(accelerate(fast)).cuda().eval()
    assert isinstance(fast.vision_model, ImageEncoderViT) and isinstance(fast.mask_decoder, MaskDecoder)
    assert isinstance(fast.prompt_encoder, PromptEncoder)
    assert type(fast).__name__ == "InternVLSAMModel" and type(fast).__module__.startswith("modeling.")
    calls = {"n": 0}
    fwd = fast.vision_model.forward

    def counting(x):
        calls["n"] += 1
        return fwd(x)

    fast.vision_model.forward = counting
    with torch.no_grad():
        out = mllm_ref.segment(fast, images.cuda(), pts.cuda(), lab.cuda(), ids.cuda(), am.cuda(), fl.cuda(),
                               multimask_output=True, reuse_embeddings=True)
    torch.cuda.synchronize()
    assert calls["n"] == 1, "the encoder must run once: outputs.image_embeddings is reused for the mask decoder"
    rel = ((out["emb"].cpu().double() - ref["emb"].double()).norm() / ref["emb"].double().norm()).item()
    rel_d = ((out["dense_map"].cpu().double() - ref["dense_map"].double()).norm() / ref["dense_map"].double().norm()).item()
    iou = _iou(out["masks"].cpu(), ref["masks"])
    err = (out["iou"].cpu() - ref["iou"]).abs().max().item()
    print(f"InternVLSAMModel (accelerated vs stock): embedding rel-L2 {rel:.3e}, LLM dense map rel-L2 {rel_d:.3e}, "
          f"multimask binarised IoU {iou:.4f}, iou-head err {err:.2e}")
    assert out["masks"].shape == (P, 3, 256, 256)
    assert rel <= 1e-2 and iou >= 0.99 and err < 3e-2
