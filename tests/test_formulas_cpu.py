"""Closed-form pieces of the CUDA path that can be checked without a GPU: the constants are read from the kernel source,
so the accuracy the comments and DESIGN.md claim is the accuracy of what is compiled."""
import os
import re

import numpy as np
from scipy.special import erf

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ullsam_b200", "csrc")


def _gelu_logistic_constants():
    src = open(os.path.join(CSRC, "common.cuh")).read()
    body = src[src.index("ULLSAM_DEVICE float gelu_logistic(float x)"):]
    body = body[: body.index("}")]
    clamp = float(re.search(r"fminf\(x \* x, ([0-9.]+)f\)", body).group(1))
    k2, k1 = map(float, re.search(r"fmaf\(t, ([-0-9.e]+)f, ([-0-9.e]+)f\)", body).groups())
    k0 = float(re.search(r"fmaf\(q, t, ([-0-9.e]+)f\)", body).group(1))
    return clamp, k0, k1, k2


def test_gelu_logistic_is_the_erf_gelu_to_2p6e_5():
    """MLPBlock's activation is nn.GELU() = x * Phi(x) (common.py:18-26).  lin1's epilogue evaluates it as
    x / (1 + 2^(x * q(min(x^2, clamp)))): the fit must stay within 2.6e-5 of the exact function on the whole real line
    (fp32 arithmetic, as in the kernel), keep q > 0 up to the clamp (so large |x| saturate the right way) and return x / -0
    for large arguments."""
    clamp, k0, k1, k2 = _gelu_logistic_constants()
    x = np.concatenate([np.linspace(-50, 50, 2_000_001), np.random.default_rng(0).normal(0, 2, 1_000_000)]).astype(np.float32)
    t = np.minimum(x * x, np.float32(clamp))
    q = (np.float32(k2) * t + np.float32(k1)) * t + np.float32(k0)
    with np.errstate(over="ignore"):
        e = np.exp2((x * q).astype(np.float32)).astype(np.float32)
    y = (x * (np.float32(1) / (np.float32(1) + e))).astype(np.float32)
    ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
    err = np.abs(y - ref).max()
    print(f"gelu_logistic: max |error| {err:.3e} (clamp {clamp}, k = {k0}, {k1}, {k2})")
    assert err <= 2.6e-5
    tt = np.linspace(0, clamp, 10001)
    qq = (k2 * tt + k1) * tt + k0
    assert (qq < 0).all() and (np.diff(qq) < 0).all(), "x * q must be monotone in x on the clamped range"
    assert y[x == 50].item() == 50.0 and y[x == -50].item() == 0.0


def test_layernorm_fold_identity():
    """LN(x) W^T + b = rstd * (x (gamma*W)^T - mean * s) + c with s = row sums of gamma*W and c = W beta + b — the identity the
    folded epilogues (EPI_QKV_FOLD / EPI_GELU_FOLD, gemm.h) rest on, including the single-pass variance from the partial
    sums the producer epilogues write."""
    rng = np.random.default_rng(1)
    M, D, N = 64, 768, 96
    x = rng.normal(0.3, 1.5, (M, D))
    g, be = 1 + 0.3 * rng.normal(size=D), 0.2 * rng.normal(size=D)
    W, b = 0.05 * rng.normal(size=(N, D)), 0.1 * rng.normal(size=N)
    mu, var = x.mean(1, keepdims=True), x.var(1, keepdims=True)
    ref = ((x - mu) / np.sqrt(var + 1e-6) * g + be) @ W.T + b
    parts = x.reshape(M, D // 128, 128)
    s1, s2 = parts.sum(-1).sum(-1), (parts * parts).sum(-1).sum(-1)          # the 128-column slices of row_stats
    mean = s1 / D
    rstd = 1 / np.sqrt(np.maximum(s2 / D - mean * mean, 0) + 1e-6)
    Wf = g * W
    got = rstd[:, None] * (x @ Wf.T - mean[:, None] * Wf.sum(1)) + (W @ be + b)
    assert np.abs(got - ref).max() <= 1e-9
