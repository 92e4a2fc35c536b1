"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol the header
declares, fails loudly without a GPU (no fallback), and the module shims keep the reference's state_dict contract."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ullsam_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return _lib.load()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ullsam.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ullsam_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from ullsam_b200 import _lib

    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ullsam.h but not exported"
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)
    assert lib.ullsam_abi_version() == 1


def test_config_struct_matches_header(lib):
    from ullsam_b200 import _lib

    # 10 leading int32 + 8 indexes + 1 float + 7 int32
    assert C.sizeof(_lib.Config) == 4 * (10 + 8 + 1 + 7)
    assert lib.ullsam_profile_categories() == 13
    names = [lib.ullsam_profile_category_name(i).decode() for i in range(13)]
    assert "attn_global" in names and "gemm_mlp_lin1" in names


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback(lib):
    from ullsam_b200 import _lib
    from ullsam_b200.engine import Engine, make_config

    h = C.c_void_p()
    cfg = make_config()
    rc = lib.ullsam_create(C.byref(cfg), 0, C.byref(h))
    assert rc == -3 and not h.value  # ULLSAM_ERR_NO_DEVICE
    assert b"no CUDA device" in lib.ullsam_last_error(None)
    with pytest.raises(_lib.UllsamError, match="no CPU fallback"):
        Engine(model="vit_b")
    d = _lib.GemmDesc()
    assert lib.ullsam_op_gemm(C.byref(d), None) == -3


def test_module_shims_keep_the_state_dict_contract():
    """Same keys / shapes as the reference modules, checked against the synthetic state_dict that
    was itself loaded strictly into the reference's Sam by oracle/make_golden.py."""
    from functools import partial

    from ullsam_b200 import synth
    from ullsam_b200.config import config
    from ullsam_b200.modeling import ImageEncoderViT, MaskDecoder, TwoWayTransformer

    for name in ("vit_b", "vit_l"):
        cfg = config(name)
        sd = synth.make_state_dict(cfg)
        enc = ImageEncoderViT(depth=cfg.depth, embed_dim=cfg.embed_dim, img_size=1024, mlp_ratio=4,
                              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=cfg.num_heads, patch_size=16,
                              qkv_bias=True, use_rel_pos=True, global_attn_indexes=list(cfg.global_attn_indexes),
                              window_size=14, out_chans=256)
        want = {k[len("image_encoder."):]: v.shape for k, v in sd.items() if k.startswith("image_encoder.")}
        got = {k: v.shape for k, v in enc.state_dict().items()}
        assert got == want
        assert enc.img_size == 1024 and len(enc.blocks) == cfg.depth
        if name == "vit_b":
            assert sum(p.numel() for p in enc.parameters()) == 89_670_912
    dec = MaskDecoder(num_multimask_outputs=3, transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048,
                                                                             num_heads=8),
                      transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256)
    want = {k[len("mask_decoder."):]: v.shape for k, v in sd.items() if k.startswith("mask_decoder.")}
    assert {k: v.shape for k, v in dec.state_dict().items()} == want
    assert dec.num_mask_tokens == 4 and sum(p.numel() for p in dec.parameters()) == 4_058_340
    res = dec.load_state_dict({k[len("mask_decoder."):]: v for k, v in sd.items() if k.startswith("mask_decoder.")})
    assert not res.missing_keys and not res.unexpected_keys


def test_module_shims_refuse_cpu_and_autograd():
    from ullsam_b200 import _lib
    from ullsam_b200.modeling import MaskDecoder, TwoWayTransformer

    dec = MaskDecoder(transformer_dim=256, transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048,
                                                                         num_heads=8))
    with torch.no_grad(), pytest.raises(_lib.UllsamError, match="no CPU fallback"):
        dec(torch.zeros(1, 256, 64, 64), torch.zeros(1, 256, 64, 64), torch.zeros(1, 2, 256),
            torch.zeros(1, 256, 64, 64), False)
    # grad mode on + trainable parameters (the reference's training loop, train_joint_v2.py:1015-1063: the inputs never
    # require grad, the parameters do): a backward would be expected, so the forward-only shim must refuse
    with pytest.raises(_lib.UllsamError, match="forward pass only"):
        dec(torch.zeros(1, 256, 64, 64), torch.zeros(1, 256, 64, 64), torch.zeros(1, 2, 256),
            torch.zeros(1, 256, 64, 64), False)
    dec.requires_grad_(False)
    x = torch.zeros(1, 256, 64, 64, requires_grad=True)
    with pytest.raises(_lib.UllsamError, match="forward pass only"):
        dec(x, torch.zeros(1, 256, 64, 64), torch.zeros(1, 2, 256), torch.zeros(1, 256, 64, 64), False)
    # frozen module + inputs without grad: passes the autograd check even with grad mode on (and then fails on the device)
    with pytest.raises(_lib.UllsamError, match="no CPU fallback"):
        dec(torch.zeros(1, 256, 64, 64), torch.zeros(1, 256, 64, 64), torch.zeros(1, 2, 256),
            torch.zeros(1, 256, 64, 64), False)
    with pytest.raises(_lib.UllsamError, match="specialised"):
        TwoWayTransformer(depth=2, embedding_dim=128, mlp_dim=2048, num_heads=8)


def test_install_substitutes_reference_classes():
    import types

    from ullsam_b200 import modeling as M

    fake = types.SimpleNamespace(ImageEncoderViT=object, MaskDecoder=object, TwoWayTransformer=object, Sam=object)
    M.install(fake)
    assert fake.ImageEncoderViT is M.ImageEncoderViT and fake.MaskDecoder is M.MaskDecoder
    assert fake.TwoWayTransformer is M.TwoWayTransformer and fake.Sam is object


def test_synthetic_inputs_are_seeded_and_shaped():
    from ullsam_b200 import synth

    a, boxes = synth.make_images(2, 256, seed=3)
    b, _ = synth.make_images(2, 256, seed=3)
    assert torch.equal(a, b) and a.shape == (2, 3, 256, 256) and a.min() >= 0 and a.max() <= 1
    assert torch.equal(a[:, 0], a[:, 1]) and boxes.shape == (2, 4)
    sp, pe, dv = synth.make_prompt_embeddings(3)
    assert sp.shape == (3, 2, 256) and pe.shape == (1, 256, 64, 64) and dv.shape == (256,)


def test_bench_flop_model_matches_baseline_md():
    import bench
    from ullsam_b200.config import config

    f = bench.flops_per_image(config("vit_b"))
    assert abs(f["total"] / 1e9 - 941.2) < 1.0          # 941.2 GF / image
    assert abs(f["attn_global"] / 4e9 - (51.54 + 0.81)) < 0.3
    fl = bench.flops_per_image(config("vit_l"))
    assert abs(fl["total"] / 1e9 - 2840.6) < 4.0


def test_prompt_encoder_shim_keeps_the_reference_state_dict_and_is_loud_on_cpu():
    """SURVEY App. B prompt_encoder.* keys; no CPU path (the shim raises instead of computing anything on the host)."""
    from ullsam_b200 import _lib
    from ullsam_b200.modeling import PromptEncoder

    pe = PromptEncoder(embed_dim=256, image_embedding_size=(64, 64), input_image_size=(1024, 1024), mask_in_chans=16)
    keys = set(pe.state_dict())
    expect = {"pe_layer.positional_encoding_gaussian_matrix", "not_a_point_embed.weight", "no_mask_embed.weight",
              "llm_scale_factor", "llm_bias"}
    expect |= {f"point_embeddings.{i}.weight" for i in range(4)}
    expect |= {f"mask_downscaling.{i}.{p}" for i in (0, 1, 3, 4, 6) for p in ("weight", "bias")}
    assert keys == expect
    assert pe.state_dict()["pe_layer.positional_encoding_gaussian_matrix"].shape == (2, 128)
    assert float(pe.llm_scale_factor) == pytest.approx(0.1) and float(pe.llm_bias) == 0.0   # prompt_encoder.py:51-52
    with pytest.raises(_lib.UllsamError):
        pe.get_dense_pe()
    with pytest.raises(_lib.UllsamError):
        pe(points=None, boxes=torch.zeros(1, 4), masks=None)


def test_postprocess_and_bridge_are_loud_on_cpu():
    from ullsam_b200 import _lib, bridge, postprocess

    with pytest.raises(_lib.UllsamError):
        postprocess.resize_masks(torch.zeros(1, 1, 256, 256), (1024, 1024))
    with pytest.raises(_lib.UllsamError):
        postprocess.mask_iou(torch.zeros(1, 1, 8, 8), torch.zeros(1, 1, 8, 8))
    mlp = torch.nn.Sequential(torch.nn.LayerNorm(1024), torch.nn.Linear(1024, 64), torch.nn.GELU(), torch.nn.Linear(64, 64))
    with pytest.raises(_lib.UllsamError):
        bridge.extract_feature(lambda x: x, mlp, torch.zeros(1, 256, 64, 64))
