"""Generate tests/golden/post.npz with the REFERENCE's own post-processing (imported from /root/reference, build container
only): Sam.postprocess_masks (modeling/sam.py:133-162, unbound on a stand-in that carries image_encoder.img_size) and
CalcIoU (train_joint_v2.py:666-696).  Pins oracle.sam_oracle.postprocess_masks / calc_iou.

    python -m oracle.make_golden_post
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("ULLSAM_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [((1024, 1024), (1024, 1024)), ((1024, 683), (1500, 1000)), ((768, 1024), (333, 444))]


def low_res(seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(2, 3, 256, 256, generator=g) * 3


@torch.no_grad()
def main():
    sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings("ignore")
    from modeling.sam import Sam  # noqa: E402
    from train_joint_v2 import CalcIoU  # noqa: E402

    stub = types.SimpleNamespace(image_encoder=types.SimpleNamespace(img_size=1024))
    out = {}
    for i, (inp, orig) in enumerate(CASES):
        m = Sam.postprocess_masks(stub, low_res(30 + i), inp, orig)
        out[f"post_{i}"] = m[:, :, ::13, ::13].numpy()
    logits = torch.nn.functional.interpolate(low_res(40), (512, 512), mode="bilinear", align_corners=False)
    target = (torch.nn.functional.interpolate(low_res(41), (512, 512), mode="bilinear", align_corners=False) > 0.5).float()
    out["iou"] = CalcIoU()(logits, target).numpy()
    path = os.path.join(ROOT, "tests", "golden", "post.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
