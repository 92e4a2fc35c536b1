"""Parameter holders shared by the module shims (reference: modeling/common.py).

These classes exist so that ``state_dict()`` / ``load_state_dict()`` / ``.parameters()`` / ``.to()`` behave exactly
like the reference's modules (same attribute tree => same keys).  They deliberately have no
eager forward: the arithmetic of the whole parent module runs in libullsam_b200.so, and there is no fallback.
"""
from __future__ import annotations

from typing import Type

import torch
import torch.nn as nn


class _NoEagerForward(nn.Module):
    def forward(self, *args, **kwargs):  # pragma: no cover - defensive
        raise RuntimeError(
            f"{type(self).__name__} is a parameter holder: only the parent module's forward is implemented "
            "(as fused sm_100a kernels); there is no eager/CPU fallback")


class MLPBlock(_NoEagerForward):
    """common.py:13-26 (lin1 -> act -> lin2)."""

    def __init__(self, embedding_dim: int, mlp_dim: int, act: Type[nn.Module] = nn.GELU) -> None:
        super().__init__()
        self.lin1 = nn.Linear(embedding_dim, mlp_dim)
        self.lin2 = nn.Linear(mlp_dim, embedding_dim)
        self.act = act()


class LayerNorm2d(_NoEagerForward):
    """common.py:31-43 (channel LayerNorm on NCHW, eps 1e-6)."""

    def __init__(self, num_channels: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps
