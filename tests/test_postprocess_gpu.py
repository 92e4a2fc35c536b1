"""Mask post-processing against the CPU oracle (which is the reference's own F.interpolate recipe):
uLLSAM's 256^2 -> 1024^2 upsampling fused with the binarisation, Sam.postprocess_masks with padding removal and odd
original sizes, CalcIoU."""
import pytest
import torch

from oracle import sam_oracle as O

pytestmark = pytest.mark.gpu


def _logits(n, c, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(n, c, h // 8, w // 8, generator=g) * 4
    return torch.nn.functional.interpolate(base, (h, w), mode="bicubic", align_corners=False) + 0.3 * torch.randn(n, c, h, w, generator=g)


def test_upsample_to_model_size_with_fused_threshold():
    from ullsam_b200 import postprocess as PP

    low = _logits(3, 1, 256, 256, 1)
    ref = O.upsample_masks(low)
    logits, binary = PP.resize_masks(low.cuda(), (1024, 1024), threshold=0.0)
    torch.cuda.synchronize()
    err = (logits.cpu() - ref).abs().max().item()
    print(f"upsample 256->1024: max|err| {err:.2e}")
    assert logits.shape == (3, 1, 1024, 1024) and err < 1e-5
    # the mask is the binarised logits except where the logit is within rounding of the threshold
    ref_b = ref > 0
    diff = (binary.cpu().bool() != ref_b) & (ref.abs() > 1e-5)
    assert binary.dtype == torch.uint8 and not diff.any()
    _, only_mask = PP.resize_masks(low.cuda(), (1024, 1024), threshold=0.0, return_logits=False)
    assert torch.equal(only_mask, binary)


@pytest.mark.parametrize("input_size,original_size", [((1024, 1024), (1024, 1024)), ((1024, 683), (1500, 1000)),
                                                      ((768, 1024), (333, 444))])
def test_postprocess_masks_like_sam(input_size, original_size):
    from ullsam_b200 import postprocess as PP

    low = _logits(2, 3, 256, 256, 2)
    ref = O.postprocess_masks(low, input_size, original_size)
    out = PP.postprocess_masks(low.cuda(), input_size, original_size)
    torch.cuda.synchronize()
    err = (out.cpu() - ref).abs().max().item()
    print(f"postprocess {input_size} -> {original_size}: max|err| {err:.2e}")
    assert out.shape == ref.shape and err < 2e-5


def test_calc_iou():
    from ullsam_b200 import postprocess as PP

    logits = O.upsample_masks(_logits(2, 3, 256, 256, 3))
    target = (O.upsample_masks(_logits(2, 3, 256, 256, 4)) > 0.5).float()
    target[1, 2] = 0           # empty target and (below) empty prediction: the smooth term decides
    logits[1, 2] = -5.0
    ref = O.calc_iou(logits, target)
    out = PP.mask_iou(logits.cuda(), target.cuda())
    torch.cuda.synchronize()
    err = (out.cpu() - ref).abs().max().item()
    print(f"CalcIoU: max|err| {err:.2e}, values {out.flatten().tolist()}")
    assert err < 1e-6
