"""Generate tests/golden/bridge.npz by calling the REFERENCE's own InternVLSAMModel methods (imported from /root/reference,
build container only) — pixel_shuffle, extract_feature's projector path and text_aware_dense_feature — as unbound methods
on a stand-in object that carries just the attributes they read (ps_version, downsample_ratio, vision_model, mlp1, mlp2),
with small seeded projectors.  The fixture pins oracle.sam_oracle's bridge functions (tests/test_oracle_golden.py).

    python -m oracle.make_golden_bridge
"""
import os
import sys
import types

import numpy as np
import torch
from torch import nn

REF = os.environ.get("ULLSAM_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_mlp(k_in, hidden, k_out, seed):
    torch.manual_seed(seed)
    m = nn.Sequential(nn.LayerNorm(k_in), nn.Linear(k_in, hidden), nn.GELU(), nn.Linear(hidden, k_out))
    with torch.no_grad():
        m[0].weight.add_(0.1 * torch.randn(k_in))
        m[0].bias.add_(0.1 * torch.randn(k_in))
    return m.eval()


@torch.no_grad()
def main():
    sys.path.insert(0, REF)
    from modeling.modeling_internvl_sam import InternVLSAMModel  # noqa: E402

    H = 64   # small LLM width keeps the fixture tiny; the functions are width-agnostic
    out = {"H": H}
    emb = torch.randn(1, 256, 64, 64, generator=torch.Generator().manual_seed(11))
    hid = torch.randn(1, 1024, H, generator=torch.Generator().manual_seed(12))
    mlp1, mlp2 = make_mlp(1024, H, H, 21), make_mlp(H, 1024, 1024, 22)
    for v in ("v2", "v1"):
        obj = types.SimpleNamespace(ps_version=v, downsample_ratio=0.5, vision_model=lambda x: x, mlp1=mlp1, mlp2=mlp2)
        obj.pixel_shuffle = types.MethodType(InternVLSAMModel.pixel_shuffle, obj)
        tokens, feats = InternVLSAMModel.extract_feature(obj, emb)
        dense = InternVLSAMModel.text_aware_dense_feature(obj, hid)
        shuf = InternVLSAMModel.pixel_shuffle(obj, emb.permute(0, 2, 3, 1), scale_factor=0.5)
        out[f"tokens_{v}"] = tokens[:, ::16].numpy()
        out[f"dense_{v}"] = dense[:, :, ::4, ::4].numpy()
        out[f"shuffle_{v}"] = shuf[:, ::4, ::4, ::8].numpy()
    path = os.path.join(ROOT, "tests", "golden", "bridge.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    import warnings
    warnings.filterwarnings("ignore")
    main()
