"""PromptEncoder drop-in (reference modeling/prompt_encoder.py) against the reference-generated
goldens and the CPU oracle: dense positional encoding, box / point sparse embeddings (labels -1 / 0 / 1, padding rule),
the LLM-conditioned dense prompt, state_dict compatibility, loud errors."""
import os

import numpy as np
import pytest
import torch

from oracle import sam_oracle as O
from ullsam_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ctx():
    from ullsam_b200.modeling import PromptEncoder

    cfg = O.config("vit_b")
    g = np.load(os.path.join(GOLD, "vit_b.npz"))
    seed = int(g["seed"])
    sd = synth.make_state_dict(cfg, seed)
    pe = PromptEncoder(embed_dim=256, image_embedding_size=(64, 64), input_image_size=(1024, 1024), mask_in_chans=16)
    sub = {k[len("prompt_encoder."):]: v for k, v in sd.items() if k.startswith("prompt_encoder.")}
    assert set(sub) == set(pe.state_dict())   # same keys as the reference module
    pe.load_state_dict(sub, strict=True)
    return dict(cfg=cfg, g=g, seed=seed, sd=sd, pe=pe.cuda().eval())


def _close(a, b, tol, what):
    err = (a.detach().float().cpu() - torch.as_tensor(b).float()).abs().max().item()
    print(f"{what}: max|err| {err:.2e}")
    assert err <= tol, what


def test_dense_pe_vs_reference_golden_and_oracle(ctx):
    pe = ctx["pe"].get_dense_pe()
    assert pe.shape == (1, 256, 64, 64) and pe.dtype == torch.float32
    _close(pe[0, :, ::8, ::8], ctx["g"]["image_pe"], 2e-5, "dense pe vs reference golden")
    _close(pe, O.dense_pe(ctx["sd"], ctx["cfg"]), 2e-5, "dense pe vs oracle")
    assert ctx["pe"].get_dense_pe() is pe   # cached


def test_box_prompt_vs_reference_golden(ctx):
    _, boxes = synth.make_images(int(ctx["g"]["B"]), 1024, seed=ctx["seed"] + 1)
    with torch.no_grad():
        sparse, dense = ctx["pe"](points=None, boxes=boxes[-1:].cuda(), masks=None)
    _close(sparse, ctx["g"]["box_sparse"], 2e-5, "box corners vs reference golden")
    assert dense.shape == (1, 256, 64, 64)
    _close(dense, O.dense_no_mask(ctx["sd"], 1, ctx["cfg"]), 0.0, "no-mask dense embedding")


def test_llm_conditioned_point_prompts_vs_reference_golden(ctx):
    P = 3
    coords, labels = synth.make_points(P, 1024, seed=ctx["seed"] + 2)
    llm = synth.make_llm_map(1, seed=ctx["seed"] + 3, C=256, g=64).repeat(P, 1, 1, 1)
    with torch.no_grad():
        sparse, dense = ctx["pe"](points=(coords.cuda(), labels.cuda()), boxes=None, masks=None,
                                  llm_hidden_states=llm.cuda())
    assert sparse.shape == (P, 5, 256) and dense.shape == (P, 256, 64, 64)
    _close(sparse, ctx["g"]["llm_sparse"], 2e-5, "point prompts (+ padding point) vs reference golden")
    _close(dense[:, :, ::8, ::8], ctx["g"]["llm_dense"], 2e-5, "LLM-conditioned dense prompt vs reference golden")
    with torch.no_grad():   # bf16 hidden states, as the app feeds them
        _, dense_bf = ctx["pe"](points=(coords.cuda(), labels.cuda()), boxes=None, masks=None,
                                llm_hidden_states=llm.cuda().bfloat16())
    _close(dense_bf, O.dense_from_llm(ctx["sd"], llm.bfloat16().float()), 2e-5, "bf16 LLM map vs oracle")


def test_bf16_module_llm_dense_uses_its_own_scale_and_bias(ctx):
    """The reference runs the prompt encoder in bfloat16 (app.py:343-345, train_joint_v2.py:1675-1676).  The fp32 copies
    of llm_scale_factor / llm_bias made for the kernel must both be alive when it runs: a freed temporary is handed to
    the next one by the caching allocator and the kernel would read scale == bias (round-1 advisor finding).  A non-zero
    bias that differs from the scale makes that aliasing visible."""
    import copy

    pe = copy.deepcopy(ctx["pe"]).bfloat16()
    with torch.no_grad():
        pe.llm_scale_factor.fill_(0.25)
        pe.llm_bias.fill_(-0.5)
    llm = synth.make_llm_map(2, seed=ctx["seed"] + 9, C=256, g=64)
    coords, labels = synth.make_points(2, 1024, seed=ctx["seed"] + 2)
    with torch.no_grad():
        _, dense = pe(points=(coords.cuda(), labels.cuda()), boxes=None, masks=None, llm_hidden_states=llm.cuda().bfloat16())
    x = llm.bfloat16().float()
    ref = torch.nn.functional.layer_norm(x.flatten(2).transpose(1, 2), (256,), eps=1e-5).transpose(1, 2).reshape(x.shape)
    ref = ref * 0.25 - 0.5                      # prompt_encoder.py:131-151
    _close(dense, ref, 2e-5, "bf16 module: LLM-conditioned dense prompt")


@pytest.mark.parametrize("with_box", [False, True])
def test_mixed_labels_and_padding_rule_vs_oracle(ctx, with_box):
    gen = torch.Generator().manual_seed(3)
    P, Np = 4, 6
    coords = torch.rand(P, Np, 2, generator=gen) * 1023
    labels = torch.randint(-1, 2, (P, Np), generator=gen)          # -1 not-a-point, 0 negative, 1 positive
    boxes = None
    ref = O.embed_points(ctx["sd"], coords, labels, ctx["cfg"], pad=not with_box)
    if with_box:
        xy = torch.rand(P, 2, generator=gen) * 500
        boxes = torch.cat([xy, xy + 100 + torch.rand(P, 2, generator=gen) * 400], dim=1)
        ref = torch.cat([ref, O.embed_boxes(ctx["sd"], boxes, ctx["cfg"])], dim=1)
    with torch.no_grad():
        sparse, _ = ctx["pe"](points=(coords.cuda(), labels.cuda()), boxes=None if boxes is None else boxes.cuda(), masks=None)
    assert sparse.shape == ref.shape == (P, Np + (2 if with_box else 1), 256)
    _close(sparse, ref, 2e-5, f"mixed labels, box={with_box}")


def test_feeds_the_decoder_like_the_reference_pipeline(ctx):
    """prompt encoder -> mask decoder on the device, against the oracle's prompt encoder -> decoder."""
    from ullsam_b200.engine import Engine

    cfg, sd = ctx["cfg"], ctx["sd"]
    eng = Engine(model="vit_b")
    eng.load_state_dict(sd)
    emb = torch.randn(1, 256, 64, 64, generator=torch.Generator().manual_seed(1)) * 0.5
    P = 3
    coords, labels = synth.make_points(P, 1024, seed=5)
    llm = synth.make_llm_map(1, seed=6, C=256, g=64).repeat(P, 1, 1, 1)
    with torch.no_grad():
        sparse, dense = ctx["pe"](points=(coords.cuda(), labels.cuda()), boxes=None, masks=None, llm_hidden_states=llm.cuda())
        m, s = eng.decoder_forward(emb.cuda(), ctx["pe"].get_dense_pe(), sparse, dense, multimask_output=False)
        m_ref, s_ref = O.decoder_forward(sd, emb, O.dense_pe(sd, cfg), O.embed_points(sd, coords, labels, cfg),
                                         O.dense_from_llm(sd, llm), False, cfg)
    torch.cuda.synchronize()
    inter = ((m.cpu() > 0) & (m_ref > 0)).sum().item()
    union = ((m.cpu() > 0) | (m_ref > 0)).sum().item()
    print(f"pipeline mask IoU {inter / max(union, 1):.4f}, iou-head err {(s.cpu() - s_ref).abs().max().item():.2e}")
    assert union == 0 or inter / union >= 0.99


def test_accelerate_swaps_the_prompt_encoder_of_a_built_model(ctx):
    from torch import nn
    from ullsam_b200 import modeling as M

    class RefPrompt(nn.Module):   # the reference module's attribute tree under a different class
        def __init__(self, src):
            super().__init__()
            self.embed_dim, self.input_image_size, self.image_embedding_size = src.embed_dim, src.input_image_size, src.image_embedding_size
            self.pe_layer, self.point_embeddings, self.not_a_point_embed = src.pe_layer, src.point_embeddings, src.not_a_point_embed
            self.llm_scale_factor, self.llm_bias = src.llm_scale_factor, src.llm_bias
            self.mask_downscaling, self.no_mask_embed = src.mask_downscaling, src.no_mask_embed

    class Model(nn.Module):
        def __init__(self, p):
            super().__init__()
            self.prompt_encoder = p

    model = M.accelerate(Model(RefPrompt(ctx["pe"])))
    assert isinstance(model.prompt_encoder, M.PromptEncoder) and model.prompt_encoder is not ctx["pe"]
    boxes = torch.tensor([[10.0, 20.0, 300.0, 400.0]], device="cuda")
    with torch.no_grad():
        a, _ = model.prompt_encoder(None, boxes, None)
        b, _ = ctx["pe"](None, boxes, None)
    assert torch.equal(a, b)


def test_mask_prompt_branch_vs_reference_golden_and_oracle(ctx):
    """Stock SAM's mask prompts (mask_downscaling, prompt_encoder.py:53-61, 105-108, 176-177) as one fused kernel: against
    the reference-generated golden (oracle/make_golden_prompt_mask.py) and the oracle, then through the CUDA decoder with
    multimask output (the Sam.forward combination of a box and a mask prompt, sam.py:101-117)."""
    from oracle.make_golden_prompt_mask import make_mask_inputs
    from ullsam_b200.engine import Engine

    g = np.load(os.path.join(GOLD, "prompt_mask.npz"))
    seed, P = int(g["seed"]), int(g["P"])
    assert seed == ctx["seed"]
    masks_in = make_mask_inputs(P, seed + 21)
    _, boxes = synth.make_images(P, 1024, seed=seed + 1)
    with torch.no_grad():
        sparse, dense = ctx["pe"](points=None, boxes=boxes.cuda(), masks=masks_in.cuda())
    assert dense.shape == (P, 256, 64, 64) and sparse.shape == (P, 2, 256)
    _close(dense[:, :, ::4, ::4], g["dense"], 2e-5, "mask-prompt dense embedding vs reference golden")
    _close(dense, O.embed_masks(ctx["sd"], masks_in), 2e-5, "mask-prompt dense embedding vs oracle")
    eng = Engine(model="vit_b")
    eng.load_state_dict(ctx["sd"])
    emb = torch.randn(1, 256, 64, 64, generator=torch.Generator().manual_seed(seed + 22)) * 0.5
    with torch.no_grad():
        m, s = eng.decoder_forward(emb.cuda(), ctx["pe"].get_dense_pe(), sparse, dense, multimask_output=True)
    ref = torch.from_numpy(g["masks"])
    a, b = m[:, :, ::4, ::4].cpu() > 0, ref > 0
    iou = (a & b).sum().item() / max((a | b).sum().item(), 1)
    print(f"mask-prompt decoder: binarised IoU vs reference golden {iou:.4f}, iou-head err {(s.cpu() - torch.from_numpy(g['iou'])).abs().max().item():.2e}")
    assert iou >= 0.99
    eng.close()


def test_errors_are_loud(ctx):
    from ullsam_b200 import _lib

    with torch.no_grad(), pytest.raises(_lib.UllsamError, match="mask prompts must be"):
        ctx["pe"](points=None, boxes=None, masks=torch.zeros(1, 1, 128, 128, device="cuda"))   # wrong mask_input_size
    import copy
    trainable = copy.deepcopy(ctx["pe"]).requires_grad_(True)
    with pytest.raises(_lib.UllsamError, match="forward-only"):      # grad mode on + trainable parameters
        trainable(points=None, boxes=torch.tensor([[1.0, 2.0, 30.0, 40.0]], device="cuda"), masks=None)
    from ullsam_b200.modeling import PromptEncoder
    cpu = PromptEncoder(256, (64, 64), (1024, 1024), 16)
    with pytest.raises(_lib.UllsamError):
        cpu.get_dense_pe()
