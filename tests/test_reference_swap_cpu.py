"""Guards for the drop-in boundary against the REAL reference classes, on CPU.

The reference sources are test infrastructure staged by oracle/make_ref.py into oracle/_ref/ (git-ignored; built by
__graft_entry__.build() where /root/reference exists, and shipped to the GPU box with the snapshot).  These tests skip —
loudly, with the reason — only when neither /root/reference nor a staged copy is available.

* class substitution: ``install(modeling)`` before ``import build_sam`` makes the reference's own ``_build_sam``
  (build_sam.py:55-106) construct the CUDA-path modules; every one of the 316 state_dict keys and shapes must be what
  the stock reference produces, and a stock state_dict must load with strict=True — for vit_b and vit_l;
* vit_h (80-wide heads) is NOT supported by the kernels: constructing it must fail loudly, not at the first forward;
* config 5: the oracle's SAM branch reproduces the real InternVLSAMModel's multimask output from the stored dense map.
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _staged_reference():
    from oracle import make_ref

    make_ref.stage(verbose=False)       # no-op on the GPU box (no /root/reference): uses what build() staged
    if not make_ref.available():
        pytest.skip("oracle/_ref not staged and /root/reference absent: run `python -m oracle.make_ref` in the build container")
    return make_ref


def _fresh_import(make_ref):
    """(modeling, build_sam) imported afresh from oracle/_ref — class substitution has to happen before build_sam binds
    the names (build_sam.py:11), so the modules are dropped from sys.modules first."""
    for name in [n for n in sys.modules if n == "build_sam" or n == "modeling" or n.startswith("modeling.")]:
        del sys.modules[name]
    return make_ref.import_reference()


@pytest.mark.parametrize("model", ["vit_b", "vit_l"])
def test_install_then_build_sam_has_the_reference_state_dict(model):
    make_ref = _staged_reference()
    from ullsam_b200 import modeling as ours

    ref_modeling, ref_build = _fresh_import(make_ref)
    torch.manual_seed(0)
    stock = ref_build.sam_model_registry[model](None)
    stock_sd = stock.state_dict()
    assert type(stock.image_encoder).__module__.startswith("modeling.")

    swapped_modeling, _ = _fresh_import(make_ref)
    ours.install(swapped_modeling)
    sys.modules.pop("build_sam", None)                 # import_reference() had already bound the stock names in it
    swapped_build = importlib.import_module("build_sam")
    torch.manual_seed(0)
    sam = swapped_build.sam_model_registry[model](None)
    assert isinstance(sam.image_encoder, ours.ImageEncoderViT) and isinstance(sam.mask_decoder, ours.MaskDecoder)
    assert isinstance(sam.mask_decoder.transformer, ours.TwoWayTransformer) and isinstance(sam.prompt_encoder, ours.PromptEncoder)
    assert type(sam).__module__.startswith("modeling.")      # Sam itself stays the reference's class (sam.py)
    sd = sam.state_dict()
    assert list(sd.keys()) == list(stock_sd.keys())           # same keys, same order
    if model == "vit_b":
        assert len(sd) == 316
    for k in sd:
        assert sd[k].shape == stock_sd[k].shape and sd[k].dtype == stock_sd[k].dtype, k
    res = sam.load_state_dict(stock_sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert sam.image_encoder.img_size == 1024                  # read at sam.py:156,171, train_joint_v2.py:1072
    # the swapped model refuses to run anywhere but on an sm_100 GPU (no CPU fallback)
    from ullsam_b200 import _lib
    with torch.no_grad(), pytest.raises(_lib.UllsamError):
        sam.image_encoder(torch.zeros(1, 3, 1024, 1024))
    _fresh_import(make_ref)                                    # leave un-substituted modules behind for other tests


def test_vit_h_is_rejected_at_construction():
    """vit_h has 1280/16 = 80-wide heads; the attention kernels are specialised for 64-wide heads.  The drop-in must say
    so when the model is BUILT (build_sam.py:14-21), not silently accept it and fail at the first forward."""
    make_ref = _staged_reference()
    from ullsam_b200 import _lib, modeling as ours

    swapped_modeling, _ = _fresh_import(make_ref)
    ours.install(swapped_modeling)
    sys.modules.pop("build_sam", None)
    swapped_build = importlib.import_module("build_sam")
    with pytest.raises(_lib.UllsamError, match="embed_dim/num_heads"):
        swapped_build.sam_model_registry["vit_h"](None)
    with pytest.raises(KeyError):
        from ullsam_b200.config import config
        config("vit_h")
    _fresh_import(make_ref)


def test_oracle_sam_branch_vs_real_internvl_sam_model_golden():
    """Config 5: P = 3 point-prompt instances, LLM-conditioned dense prompt from the real InternVLSAMModel
    (oracle/make_golden_mllm.py), multimask and single-mask output.  Pins the oracle's prompt-encoder + decoder on the
    exact tensors the full model produces."""
    from oracle import sam_oracle as O
    from ullsam_b200 import synth

    g = np.load(os.path.join(GOLD, "mllm.npz"))
    seed, P = int(g["seed"]), int(g["P"])
    cfg = O.config("vit_b")
    sd = synth.make_state_dict(cfg, seed)
    images, _ = synth.make_images(1, 1024, seed=seed + 11)
    pts, lab = synth.make_points(P, 1024, seed=seed + 12)
    dense_map = torch.from_numpy(g["dense_map_f16"]).float()
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        emb = O.encoder_forward(sd, images, cfg)
        assert np.abs(emb[:, :, ::4, ::4].numpy() - g["emb"]).max() < 2e-4
        pe = O.dense_pe(sd, cfg)
        sparse = O.embed_points(sd, pts, lab, cfg)
        dense = O.dense_from_llm(sd, dense_map.repeat(P, 1, 1, 1))
        for tag, mm in (("multi", True), ("single", False)):
            masks, iou = O.decoder_forward(sd, emb, pe, sparse, dense, mm, cfg)
            assert masks.shape == (P, 3 if mm else 1, 256, 256)
            err = np.abs(masks[:, :, ::4, ::4].numpy() - g[f"masks_{tag}"]).max()
            print(f"{tag}: mask logits max|err| {err:.2e}, iou err {np.abs(iou.numpy() - g[f'iou_{tag}']).max():.2e}")
            assert err < 2e-3 and np.abs(iou.numpy() - g[f"iou_{tag}"]).max() < 2e-4
