"""Hyper-parameters of the SAM path (what build_sam._build_sam fixes, build_sam.py:14-98)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple


@dataclass
class SamConfig:
    """Hyper-parameters fixed by build_sam._build_sam (build_sam.py:55-98)."""

    name: str = "vit_b"
    img_size: int = 1024
    patch_size: int = 16
    in_chans: int = 3
    embed_dim: int = 768
    depth: int = 12
    num_heads: int = 12
    mlp_ratio: float = 4.0
    out_chans: int = 256
    window_size: int = 14
    global_attn_indexes: Tuple[int, ...] = (2, 5, 8, 11)
    enc_ln_eps: float = 1e-6
    # decoder
    transformer_dim: int = 256
    dec_depth: int = 2
    dec_heads: int = 8
    dec_mlp_dim: int = 2048
    num_multimask_outputs: int = 3
    iou_head_depth: int = 3
    iou_head_hidden_dim: int = 256
    mask_in_chans: int = 16
    extra: dict = field(default_factory=dict)

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def num_mask_tokens(self) -> int:
        return self.num_multimask_outputs + 1


def config(name: str) -> SamConfig:
    """build_sam.py:24-44 (vit_l / vit_b) plus a tiny configuration used to pin the oracle cheaply.  vit_h
    (build_sam.py:14-21: 1280 / 16 heads = 80-wide heads) is NOT supported: the attention kernels are specialised for
    64-wide heads, and uLLSAM itself only ships ViT-B (build_sam.py:50-52 registry, train_joint_v2.py:1365)."""
    if name == "vit_b":
        return SamConfig()
    if name == "vit_l":
        return SamConfig(name="vit_l", embed_dim=1024, depth=24, num_heads=16, global_attn_indexes=(5, 11, 17, 23))
    if name == "tiny":
        # 20x20 token grid: windows of 14 need padding to 28 (edge windows are partly padding), one global block.
        return SamConfig(name="tiny", img_size=320, embed_dim=128, depth=3, num_heads=2, global_attn_indexes=(1,),
                         out_chans=256)
    raise KeyError(name)
