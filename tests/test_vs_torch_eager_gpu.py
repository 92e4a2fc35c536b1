"""SURVEY.md §8(d) "second comparison": the same forward as plain PyTorch eager ops in bf16 autocast on the SAME B200
(the recipe the reference uses on GPUs, train_joint_v2.py:1665,1675-1676), next to the sm_100a path.

/root/reference does not exist on the GPU box, so the eager arm is the oracle's functional restatement of the reference
modules (oracle/sam_oracle.py — identical torch ops, weights on the device).  The test prints both throughputs; it
asserts only that the hand-written path is not slower (the numbers are for profiles/, not a pass/fail bar)."""
import os

import pytest
import torch

from oracle import sam_oracle as O
from ullsam_b200 import synth

pytestmark = pytest.mark.gpu


def _time(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def test_cuda_path_vs_pytorch_eager_bf16_on_this_gpu():
    from ullsam_b200.engine import Engine

    cfg = O.config("vit_b")
    sd = synth.make_state_dict(cfg, 1234)
    eng = Engine(model="vit_b")
    eng.load_state_dict(sd)
    B = 8
    images, _ = synth.make_images(B, 1024, seed=7)
    images = images.cuda()
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    pe = O.dense_pe(sd_gpu, cfg)
    sparse = torch.randn(B, 2, 256, device="cuda") * 0.5
    dense = sd_gpu["prompt_encoder.no_mask_embed.weight"].reshape(1, 256, 1, 1).expand(1, 256, 64, 64).contiguous()

    def eager():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            emb = O.encoder_forward(sd_gpu, images, cfg)
            outs = [O.decoder_forward(sd_gpu, emb[i:i + 1].float(), pe, sparse[i:i + 1], dense, False, cfg) for i in range(B)]
        return emb, outs

    def ours():
        emb = eng.encoder_forward(images)
        return emb, eng.decoder_forward(emb, pe, sparse, dense.expand(B, 256, 64, 64).contiguous(), multimask_output=False)

    ms_eager = _time(eager, 2)
    ms_ours = _time(ours, 5)
    emb_e, _ = eager()
    emb_o, _ = ours()
    rel = ((emb_e.float() - emb_o).norm() / emb_o.norm()).item()
    print(f"PyTorch eager bf16 autocast: {B / ms_eager * 1e3:8.1f} images/s ({ms_eager:.1f} ms / {B} images)")
    print(f"ullsam_b200 (sm_100a)      : {B / ms_ours * 1e3:8.1f} images/s ({ms_ours:.1f} ms / {B} images)   "
          f"speed-up {ms_eager / ms_ours:.1f}x, embeddings differ by rel-L2 {rel:.2e}")
    assert ms_ours < ms_eager
    assert rel < 3e-2   # two reduced-precision paths against each other (each is within 1e-2 of fp32)


def test_fused_postprocess_vs_pytorch_eager_on_this_gpu():
    """Downstream step: 32 masks 256^2 -> binary 1024^2, fused kernel vs F.interpolate + threshold."""
    from ullsam_b200 import postprocess as PP

    low = torch.randn(32, 1, 256, 256, device="cuda")

    def eager():
        return torch.nn.functional.interpolate(low, (1024, 1024), mode="bilinear", align_corners=False) > 0

    def ours():
        return PP.resize_masks(low, (1024, 1024), threshold=0.0, return_logits=False)[1]

    ms_e, ms_o = _time(eager, 20), _time(ours, 20)
    same = (eager() == ours().bool()).float().mean().item()
    print(f"post-processing of 32 masks: PyTorch eager {ms_e * 1e3:.0f} us, fused kernel {ms_o * 1e3:.0f} us "
          f"({ms_e / ms_o:.1f}x), identical pixels {same:.6f}")
    assert same > 0.99999
