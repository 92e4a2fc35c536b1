"""ctypes binding of libullsam_b200.so (the C ABI declared in include/ullsam.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C ullsam_b200/csrc``.  There is no
fallback: if the shared object is missing, ``load()`` raises, and every compute entry point returns
``ULLSAM_ERR_NO_DEVICE`` on a machine without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ULLSAM_B200_LIB") or os.path.join(_HERE, "lib", "libullsam_b200.so")

OK = 0
F32, BF16, F16 = 0, 1, 2

EPI_BF16, EPI_RES_F32, EPI_QKV, EPI_PATCH, EPI_LN_BF16, EPI_LN_NCHW, EPI_F32 = range(7)
EPI_LN_RES = 7   # x = acc + bias + residual; LayerNorm(256) -> out fp32 (may alias residual) and out2 bf16 (decoder norm4)
EPI_RES_STATS, EPI_PATCH_STATS, EPI_QKV_FOLD, EPI_GELU_FOLD = 12, 13, 14, 15  # LayerNorm folded into the GEMMs around it
EPI_RES_RED = 16  # EPI_RES_F32 in place, the add done by TMA reduce-add stores
STAT_SLOTS = 8   # 128-column slices per row of the row_stats buffer
EPI_RES_LN = 11  # +residual fp32 and the LayerNorm that follows it -> bf16 (optionally window-partitioned)
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2


class UllsamError(RuntimeError):
    pass


class Config(C.Structure):
    """Mirror of ``ullsam_config`` (include/ullsam.h)."""

    _fields_ = [
        ("img_size", C.c_int32),
        ("patch_size", C.c_int32),
        ("in_chans", C.c_int32),
        ("embed_dim", C.c_int32),
        ("depth", C.c_int32),
        ("num_heads", C.c_int32),
        ("mlp_dim", C.c_int32),
        ("out_chans", C.c_int32),
        ("window_size", C.c_int32),
        ("num_global", C.c_int32),
        ("global_attn_indexes", C.c_int32 * 8),
        ("enc_ln_eps", C.c_float),
        ("transformer_dim", C.c_int32),
        ("dec_depth", C.c_int32),
        ("dec_heads", C.c_int32),
        ("dec_mlp_dim", C.c_int32),
        ("num_mask_tokens", C.c_int32),
        ("iou_head_hidden", C.c_int32),
        ("iou_head_depth", C.c_int32),
    ]


class GemmDesc(C.Structure):
    """Mirror of ``ullsam_gemm_desc`` (include/ullsam.h)."""

    _fields_ = [
        ("epi", C.c_int32),
        ("M", C.c_int32),
        ("N", C.c_int32),
        ("K", C.c_int32),
        ("a_conv3", C.c_int32),
        ("conv_c", C.c_int32),
        ("A", C.c_void_p),
        ("W", C.c_void_p),
        ("bias", C.c_void_p),
        ("out", C.c_void_p),
        ("ldc", C.c_int32),
        ("act", C.c_int32),
        ("residual", C.c_void_p),
        ("pos", C.c_void_p),
        ("pos_rows", C.c_int32),
        ("q", C.c_void_p),
        ("k", C.c_void_p),
        ("v", C.c_void_p),
        ("T", C.c_int32),
        ("heads", C.c_int32),
        ("ln_w", C.c_void_p),
        ("ln_b", C.c_void_p),
        ("ln_eps", C.c_float),
        ("hw", C.c_int32),
        ("out2", C.c_void_p),
        ("ln_partition", C.c_int32),
        ("row_stats", C.c_void_p),
        ("fold_s", C.c_void_p),
        ("pad_bias", C.c_void_p),
        ("kmax_sq", C.c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/ullsam.h declares.
PROTOTYPES = {
    "ullsam_abi_version": (C.c_int, []),
    "ullsam_last_error": (C.c_char_p, [C.c_void_p]),
    "ullsam_create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(C.c_void_p)]),
    "ullsam_destroy": (None, [C.c_void_p]),
    "ullsam_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_int]),
    "ullsam_set_weight_async": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.c_int,
                                          C.c_void_p]),
    "ullsam_missing_weights": (C.c_int, [C.c_void_p, C.c_int]),
    "ullsam_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "ullsam_encoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "ullsam_decoder_forward": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
         C.c_int, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "ullsam_decoder_predict_masks": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
         C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "ullsam_transformer_forward": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "ullsam_segment_host": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
         C.c_void_p, C.c_void_p],
    ),
    "ullsam_launch_count": (C.c_ulonglong, [C.c_void_p]),
    "ullsam_profile_begin": (C.c_int, [C.c_void_p]),
    "ullsam_profile_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]),
    "ullsam_profile_categories": (C.c_int, []),
    "ullsam_profile_category_name": (C.c_char_p, [C.c_int]),
    "ullsam_op_gemm": (C.c_int, [C.POINTER(GemmDesc), C.c_void_p]),
    "ullsam_op_layernorm": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p],
    ),
    "ullsam_op_attention": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p],
    ),
    "ullsam_op_im2col": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ullsam_op_cast_bf16": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ullsam_op_pixel_shuffle": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ullsam_op_resize_masks": (
        C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                  C.c_void_p]),
    "ullsam_op_mask_iou": (
        C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "ullsam_op_prompt_dense_pe": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ullsam_op_prompt_sparse": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
         C.c_float, C.c_float, C.c_void_p, C.c_void_p],
    ),
    "ullsam_op_prompt_mask_embed": (
        C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "ullsam_op_prompt_llm_dense": (
        C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UllsamError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C ullsam_b200/csrc`).  There is no CPU/eager fallback for this path."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error(handle: Optional[int] = None) -> str:
    msg = load().ullsam_last_error(C.c_void_p(handle) if handle else None)
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, handle: Optional[int] = None, what: str = "") -> None:
    if status != OK:
        raise UllsamError(f"{what or 'ullsam call'} failed (status {status}): {last_error(handle)}")


def ptr(t) -> Optional[int]:
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream_ptr(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream
