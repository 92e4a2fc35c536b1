"""Generate tests/golden/prompt_mask.npz: stock SAM's mask-prompt branch (PromptEncoder._embed_masks ==
mask_downscaling, modeling/prompt_encoder.py:53-61, 105-108, 176-177) from the REFERENCE's own PromptEncoder
(imported from oracle/_ref, staged from /root/reference), on the seeded synthetic weights.  Also stores the masks and
IoU predictions of the reference decoder fed with that dense embedding plus a box prompt (the Sam.forward combination,
sam.py:101-117).  Pins oracle.sam_oracle.embed_masks.

    python -m oracle.make_golden_prompt_mask
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 1234


def make_mask_inputs(P: int, seed: int) -> torch.Tensor:
    """Low-res mask logits as Sam.forward receives them ([P, 1, 256, 256]): smooth blobs, a few units in magnitude."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.randn(P, 1, 16, 16, generator=g) * 4
    return torch.nn.functional.interpolate(coarse, (256, 256), mode="bilinear", align_corners=False) + \
        0.3 * torch.randn(P, 1, 256, 256, generator=g)


@torch.no_grad()
def main():
    from oracle import make_ref, sam_oracle as O
    from ullsam_b200 import synth

    make_ref.stage()
    _, build_sam = make_ref.import_reference()
    cfg = O.config("vit_b")
    sd = synth.make_state_dict(cfg, SEED)
    torch.manual_seed(0)
    sam = build_sam.sam_model_registry["vit_b"](None)
    sam.load_state_dict(sd, strict=False)
    P = 2
    masks_in = make_mask_inputs(P, SEED + 21)
    _, boxes = synth.make_images(P, 1024, seed=SEED + 1)
    sparse, dense = sam.prompt_encoder(points=None, boxes=boxes, masks=masks_in)
    emb = torch.randn(1, 256, 64, 64, generator=torch.Generator().manual_seed(SEED + 22)) * 0.5
    m, s = sam.mask_decoder(image_embeddings=emb, image_pe=sam.prompt_encoder.get_dense_pe(), sparse_prompt_embeddings=sparse,
                            dense_prompt_embeddings=dense, multimask_output=True)
    out = {"seed": SEED, "P": P, "dense": dense[:, :, ::4, ::4].numpy(), "dense_l2": dense.pow(2).sum().sqrt().item(),
           "sparse": sparse.numpy(), "masks": m[:, :, ::4, ::4].numpy(), "iou": s.numpy()}
    path = os.path.join(ROOT, "tests", "golden", "prompt_mask.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    warnings.filterwarnings("ignore")
    main()
