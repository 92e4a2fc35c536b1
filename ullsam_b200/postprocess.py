"""Mask post-processing on the device: the step immediately downstream of the mask decoder.

* :func:`resize_masks` — ``F.interpolate(..., mode="bilinear", align_corners=False)`` fused with the binarisation, so the
  upsampled fp32 logits (4 MB per 1024x1024 mask) need not be written at all when only the mask is wanted;
* :func:`postprocess_masks` — drop-in for ``Sam.postprocess_masks`` (sam.py:133-162);
* :func:`mask_iou` — ``CalcIoU`` (train_joint_v2.py:666-696).

All three call ``libullsam_b200.so`` (``ullsam_op_resize_masks`` / ``ullsam_op_mask_iou``); there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from .ops import _on_tensor_device, _require_cuda, check, current_stream_ptr, ptr


@_on_tensor_device
def resize_masks(masks: torch.Tensor, size: Tuple[int, int], crop: Optional[Tuple[int, int]] = None,
                 threshold: Optional[float] = None, return_logits: bool = True
                 ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """masks [B, C, h, w] -> (fp32 logits [B, C, H, W] or None, uint8 0/1 mask or None).

    ``crop=(ch, cw)`` resamples only ``masks[..., :ch, :cw]`` (the padding removal of sam.py:160)."""
    _require_cuda(masks)
    x = masks.float().contiguous()
    B, Cc, h, w = x.shape
    ch, cw = crop if crop is not None else (h, w)
    H, W = int(size[0]), int(size[1])
    out_f = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device) if return_logits else None
    out_b = torch.empty((B, Cc, H, W), dtype=torch.uint8, device=x.device) if threshold is not None else None
    if out_f is None and out_b is None:
        raise _lib.UllsamError("resize_masks: ask for the logits, a threshold, or both")
    check(_lib.load().ullsam_op_resize_masks(ptr(x), B * Cc, h, w, int(ch), int(cw), H, W,
                                             float(threshold if threshold is not None else 0.0), ptr(out_f), ptr(out_b),
                                             current_stream_ptr(x.device)), what="ullsam_op_resize_masks")
    return out_f, out_b


def postprocess_masks(masks: torch.Tensor, input_size: Tuple[int, ...], original_size: Tuple[int, ...],
                      img_size: int = 1024) -> torch.Tensor:
    """``Sam.postprocess_masks`` (sam.py:133-162): upscale to the model size, remove the padding, resize to the
    original image size.  Returns fp32 logits [B, C, *original_size]."""
    up, _ = resize_masks(masks, (img_size, img_size))
    out, _ = resize_masks(up, (int(original_size[0]), int(original_size[1])), crop=(int(input_size[0]), int(input_size[1])))
    return out


@_on_tensor_device
def mask_iou(logits: torch.Tensor, target: torch.Tensor, threshold: float = 0.0, smooth: float = 1e-7) -> torch.Tensor:
    """``CalcIoU`` (train_joint_v2.py:666-696): logits, target [B, N, H, W] -> IoU [B, N] of (sigmoid > 0.5) vs target."""
    _require_cuda(logits, target)
    a, b = logits.float().contiguous(), target.float().contiguous()
    if a.shape != b.shape:
        raise _lib.UllsamError("mask_iou: logits and target shapes differ")
    B, N = a.shape[:2]
    hw = a[0, 0].numel()
    out = torch.empty((B, N), dtype=torch.float32, device=a.device)
    check(_lib.load().ullsam_op_mask_iou(ptr(a), ptr(b), B * N, hw, float(threshold), float(smooth), ptr(out),
                                         current_stream_ptr(a.device)), what="ullsam_op_mask_iou")
    return out
