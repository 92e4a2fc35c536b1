"""Generate tests/golden/mllm.npz: config 5 of BASELINE.json — the SAM branch INSIDE the reference's real
`InternVLSAMModel` (modeling_internvl_sam.py:106-201, 242-270), multimask output, driven as train_joint_v2.train_epoch /
app.py drive it (oracle/mllm_ref.py).  Runs where /root/reference exists (it stages oracle/_ref first).

    python -m oracle.make_golden_mllm

Stored: the dense map the model hands to the prompt encoder (outputs.hidden_states after text_aware_dense_feature,
[1,256,64,64], rounded to fp16 so the fixture stays ~2 MB — the masks below are computed FROM THE ROUNDED MAP, so the
fixture is self-consistent), sub-sampled vision tokens / embeddings, and, for P = 3 point-prompt instances with multimask_output = True and False, the
mask logits (every 4th pixel), the FULL-resolution binarised masks (bit-packed — the IoU bar of north_star is defined
on them) and the IoU predictions.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED, LLM_SEED, P = 1234, 4321, 3


@torch.no_grad()
def main():
    from oracle import make_ref, mllm_ref, sam_oracle as O
    from ullsam_b200 import synth

    make_ref.stage()
    cfg = O.config("vit_b")
    sd = synth.make_state_dict(cfg, SEED)
    m = mllm_ref.build(sd, seed=LLM_SEED)
    images, _ = synth.make_images(1, 1024, seed=SEED + 11)
    pts, lab = synth.make_points(P, 1024, seed=SEED + 12)
    ids, am, fl = mllm_ref.make_inputs(1, seed=SEED + 13)
    out = m(pixel_values=images, input_ids=ids, attention_mask=am, image_flags=fl, output_hidden_states=True, return_dict=True)
    tokens, emb = m.extract_feature(images)
    dense_map = out.hidden_states.half().float()          # what the fixture stores
    pe = m.prompt_encoder.get_dense_pe()
    res = {"seed": SEED, "llm_seed": LLM_SEED, "P": P, "dense_map_f16": dense_map.half().numpy(),
           "emb": emb[:, :, ::4, ::4].numpy(), "emb_l2": emb.pow(2).sum().sqrt().item(),
           "vit_tokens": tokens[:, ::16, ::4].numpy(), "logits_l2": out.logits.float().pow(2).sum().sqrt().item()}
    sparse, dense = m.prompt_encoder(points=(pts, lab), boxes=None, masks=None, llm_hidden_states=dense_map.repeat(P, 1, 1, 1))
    for mm in (True, False):
        masks, iou = m.mask_decoder(image_embeddings=emb, image_pe=pe, sparse_prompt_embeddings=sparse,
                                    dense_prompt_embeddings=dense, multimask_output=mm)
        tag = "multi" if mm else "single"
        res[f"masks_{tag}"] = masks[:, :, ::4, ::4].numpy()
        res[f"bits_{tag}"] = np.packbits((masks > 0).numpy().reshape(-1))     # full-resolution binarised masks
        res[f"pos_frac_{tag}"] = (masks > 0).float().mean(dim=(2, 3)).numpy()
        res[f"iou_{tag}"] = iou.numpy()
    path = os.path.join(ROOT, "tests", "golden", "mllm.npz")
    np.savez_compressed(path, **res)
    print(path, os.path.getsize(path) // 1024, "KiB", {k: getattr(v, "shape", v) for k, v in res.items()})


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    warnings.filterwarnings("ignore")
    torch.set_num_threads(os.cpu_count() or 1)
    main()
