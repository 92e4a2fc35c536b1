"""MLLM bridge (reference modeling/modeling_internvl_sam.py:211-270) against the CPU oracle: the
pixel shuffle and its inverse (both ps_versions, bit-exact data movement), the two projector MLPs on the tcgen05 GEMM
(bf16 operands: relative error bar 1e-2 like the encoder), and the method patch on an InternVLSAMModel-like object."""
import pytest
import torch
from torch import nn

from oracle import sam_oracle as O

pytestmark = pytest.mark.gpu


def _mlp(k_in, hidden, k_out, seed):
    torch.manual_seed(seed)
    m = nn.Sequential(nn.LayerNorm(k_in), nn.Linear(k_in, hidden), nn.GELU(), nn.Linear(hidden, k_out))
    with torch.no_grad():
        m[0].weight.add_(0.1 * torch.randn(k_in))
        m[0].bias.add_(0.1 * torch.randn(k_in))
    return m


def _params(m):
    return {k: v.detach() for k, v in m.state_dict().items()}


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.mark.parametrize("ps_version", ["v2", "v1"])
def test_pixel_shuffle_and_inverse_are_exact(ps_version):
    from ullsam_b200 import ops

    x = torch.randn(2, 256, 64, 64)
    ref = O.pixel_shuffle(x.permute(0, 2, 3, 1), 0.5, ps_version).reshape(2, 1024, 1024)
    tok = ops.pixel_shuffle(x.cuda(), to_tokens=True, ps_version=ps_version)
    assert torch.equal(tok.cpu(), ref)
    back = ops.pixel_shuffle(tok, to_tokens=False, ps_version=ps_version)
    assert torch.equal(back.cpu(), x)   # the inverse is what text_aware_dense_feature applies after mlp2


def test_projectors_vs_oracle():
    from ullsam_b200 import bridge

    H = 2048
    mlp1, mlp2 = _mlp(1024, H, H, 1), _mlp(H, 1024, 1024, 2)
    emb = torch.randn(2, 256, 64, 64, generator=torch.Generator().manual_seed(3)) * 0.7
    ref_tokens = O.mllm_extract_feature(emb, _params(mlp1))
    tokens, feats = bridge.extract_feature(lambda x: x, mlp1.cuda(), emb.cuda())
    torch.cuda.synchronize()
    e1 = _rel(tokens.cpu().float(), ref_tokens)
    print(f"extract_feature (mlp1): rel-L2 {e1:.2e}")
    assert tokens.shape == (2, 1024, H) and feats.shape == emb.shape and e1 < 1e-2

    hidden = torch.randn(2, 1024, H, generator=torch.Generator().manual_seed(4))
    ref_map = O.text_aware_dense_feature(hidden, _params(mlp2))
    dense = bridge.text_aware_dense_feature(mlp2.cuda(), hidden.cuda())
    torch.cuda.synchronize()
    e2 = _rel(dense.cpu().float(), ref_map)
    print(f"text_aware_dense_feature (mlp2): rel-L2 {e2:.2e}")
    assert dense.shape == (2, 256, 64, 64) and e2 < 1e-2


def test_patch_on_an_internvl_sam_like_model():
    from ullsam_b200 import bridge

    class Model(nn.Module):
        def __init__(self):
            super().__init__()
            self.vision_model = nn.Identity()
            self.mlp1, self.mlp2 = _mlp(1024, 2048, 2048, 5), _mlp(2048, 1024, 1024, 6)
            self.ps_version, self.downsample_ratio = "v2", 0.5

    m = bridge.patch_internvl_sam(Model().cuda().bfloat16())
    emb = (torch.randn(1, 256, 64, 64, generator=torch.Generator().manual_seed(7)) * 0.7).cuda().bfloat16()
    tokens, feats = m.extract_feature(emb)
    dense = m.text_aware_dense_feature(tokens)
    torch.cuda.synchronize()
    assert tokens.dtype == torch.bfloat16 and tokens.shape == (1, 1024, 2048)
    assert dense.dtype == torch.bfloat16 and dense.shape == (1, 256, 64, 64) and torch.isfinite(dense.float()).all()
    ref = O.text_aware_dense_feature(O.mllm_extract_feature(emb.float().cpu(), _params(m.mlp1.float().cpu())),
                                     _params(m.mlp2.float().cpu()))
    e = _rel(dense.float().cpu(), ref)
    print(f"patched model, bf16 end to end: rel-L2 {e:.2e}")
    assert e < 2e-2
