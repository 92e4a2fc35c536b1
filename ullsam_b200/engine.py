"""Handle-level Python wrapper over the C ABI: one ``Engine`` per (device, model).

torch is used for device memory, streams and (in ``ullsam_b200.parallel``) ``torch.distributed`` only; all of
the arithmetic happens in libullsam_b200.so.  There is no fallback path: constructing an Engine without an
sm_100 GPU raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import Config, check, ptr

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}


def make_config(embed_dim: int = 768, depth: int = 12, num_heads: int = 12, global_attn_indexes: Sequence[int] = (2, 5, 8, 11),
                mlp_ratio: float = 4.0, img_size: int = 1024, patch_size: int = 16, in_chans: int = 3, out_chans: int = 256,
                window_size: int = 14, enc_ln_eps: float = 1e-6, transformer_dim: int = 256, dec_depth: int = 2,
                dec_heads: int = 8, dec_mlp_dim: int = 2048, num_mask_tokens: int = 4, iou_head_hidden: int = 256,
                iou_head_depth: int = 3) -> Config:
    c = Config()
    c.img_size, c.patch_size, c.in_chans = img_size, patch_size, in_chans
    c.embed_dim, c.depth, c.num_heads = embed_dim, depth, num_heads
    c.mlp_dim = int(embed_dim * mlp_ratio)
    c.out_chans, c.window_size = out_chans, window_size
    g = list(global_attn_indexes)
    if len(g) > 8:
        raise _lib.UllsamError("at most 8 global-attention blocks are supported")
    c.num_global = len(g)
    for i, v in enumerate(g):
        c.global_attn_indexes[i] = int(v)
    c.enc_ln_eps = enc_ln_eps
    c.transformer_dim, c.dec_depth, c.dec_heads, c.dec_mlp_dim = transformer_dim, dec_depth, dec_heads, dec_mlp_dim
    c.num_mask_tokens, c.iou_head_hidden, c.iou_head_depth = num_mask_tokens, iou_head_hidden, iou_head_depth
    return c


MODEL_CONFIGS = {
    # build_sam.py:14-44
    "vit_b": dict(embed_dim=768, depth=12, num_heads=12, global_attn_indexes=(2, 5, 8, 11)),
    "vit_l": dict(embed_dim=1024, depth=24, num_heads=16, global_attn_indexes=(5, 11, 17, 23)),
    # vit_h (1280 / 16 = 80-wide heads) is not supported: the attention kernels are built for 64-wide heads
}


class Engine:
    """Owns one ``ullsam_handle``.  All tensor arguments must live on ``device``."""

    def __init__(self, config: Optional[Config] = None, device: Optional[int] = None, model: str = "vit_b"):
        self._h: Optional[int] = None
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.UllsamError("ullsam_b200 needs a CUDA device (sm_100); there is no CPU fallback")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.config = config if config is not None else make_config(**MODEL_CONFIGS[model])
        h = C.c_void_p()
        check(self.lib.ullsam_create(C.byref(self.config), self.device_index, C.byref(h)), None, "ullsam_create")
        self._h = h.value

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if self._h:
            self.lib.ullsam_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------ weights
    def set_weight(self, key: str, tensor: torch.Tensor) -> None:
        t = tensor.detach()
        if t.dtype not in _DT:
            t = t.float()
        t = t.contiguous()
        shape = (C.c_int64 * t.dim())(*t.shape)
        if t.is_cuda and t.device == self.device:
            # stream-ordered upload on torch's current stream: no host sync per tensor, and a temporary made by the
            # .float()/.contiguous() above stays valid because torch's allocator recycles it in the same stream's order
            with torch.cuda.device(self.device):
                rc = self.lib.ullsam_set_weight_async(self._h, key.encode(), ptr(t), _DT[t.dtype], shape, t.dim(),
                                                      self._stream())
            check(rc, self._h, f"ullsam_set_weight_async({key})")
            return
        # host tensors and tensors on another GPU: torch stages them onto the engine's device, then the same path
        return self.set_weight(key, t.to(self.device))

    def load_state_dict(self, sd: Dict[str, torch.Tensor], prefixes: Iterable[Tuple[str, str]] = (("", ""),),
                        only: Optional[str] = None) -> int:
        """Feed every tensor of ``sd`` whose (re-prefixed) key belongs to the path.  ``prefixes`` maps a prefix found
        in ``sd`` to the canonical one ("image_encoder." / "mask_decoder."), e.g. uLLSAM checkpoints use
        "vision_model." for the encoder (train_joint_v2.py:1472-1475).  Returns the number of tensors loaded."""
        n = 0
        for k, v in sd.items():
            for src, dst in prefixes:
                if not k.startswith(src):
                    continue
                key = dst + k[len(src):]
                if not (key.startswith("image_encoder.") or key.startswith("mask_decoder.")):
                    continue
                if only and not key.startswith(only):
                    continue
                self.set_weight(key, v)
                n += 1
                break
        return n

    def missing(self, which: int) -> int:
        return self.lib.ullsam_missing_weights(self._h, which)

    # ------------------------------------------------------------------ instrumentation
    def launch_count(self) -> int:
        return int(self.lib.ullsam_launch_count(self._h))

    def profile_begin(self) -> None:
        check(self.lib.ullsam_profile_begin(self._h), self._h, "ullsam_profile_begin")

    def profile_end(self) -> Dict[str, Tuple[float, int]]:
        """{category: (milliseconds, marked launches)} accumulated since ``profile_begin``."""
        n = self.lib.ullsam_profile_categories()
        ms = (C.c_float * n)()
        cnt = (C.c_int * n)()
        check(self.lib.ullsam_profile_end(self._h, ms, cnt, n), self._h, "ullsam_profile_end")
        return {self.lib.ullsam_profile_category_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    # ------------------------------------------------------------------ forward passes
    def encoder_forward(self, images: torch.Tensor, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """ImageEncoderViT.forward: [B,3,1024,1024] (fp32 / bf16) -> [B,256,64,64]."""
        if not images.is_cuda or images.device != self.device:
            raise _lib.UllsamError("encoder_forward needs a CUDA tensor on the engine's device (no CPU path)")
        if images.dtype == torch.float16:
            images = images.float()
        images = images.contiguous()
        B = images.shape[0]
        if tuple(images.shape[1:]) != (3, 1024, 1024):
            raise _lib.UllsamError(f"encoder_forward expects [B,3,1024,1024], got {tuple(images.shape)}")
        out = torch.empty((B, 256, 64, 64), dtype=out_dtype, device=self.device)
        check(self.lib.ullsam_encoder_forward(self._h, ptr(images), _DT[images.dtype], B, ptr(out), _DT[out_dtype],
                                              self._stream()), self._h, "ullsam_encoder_forward")
        return out

    @staticmethod
    def _dense_args(dense: torch.Tensor, R: int):
        """Recognise PromptEncoder's broadcast no_mask_embed ([bs,256,64,64] with stride 0 over H, W)."""
        if dense.dim() == 2:
            d = dense.float().contiguous()
            return d, 1, d.shape[0]
        if dense.dim() == 4 and dense.stride(2) == 0 and dense.stride(3) == 0:
            d = dense[:, :, 0, 0]
            if dense.stride(0) == 0:
                d = d[:1]
            d = d.float().contiguous()
            return d, 1, d.shape[0]
        d = dense.float().contiguous()
        return d, 0, d.shape[0]

    def decoder_forward(self, image_embeddings: torch.Tensor, image_pe: torch.Tensor, sparse: torch.Tensor,
                        dense: torch.Tensor, multimask_output: Optional[bool], prompts_per_image: Optional[int] = None
                        ) -> Tuple[torch.Tensor, torch.Tensor]:
        """MaskDecoder.forward (``multimask_output`` bool) or predict_masks (``multimask_output is None``)."""
        emb = image_embeddings.float().contiguous()
        n_img = emb.shape[0]
        R = sparse.shape[0]
        P = prompts_per_image if prompts_per_image is not None else max(R // n_img, 1)
        if n_img * P != R:
            raise _lib.UllsamError(f"decoder: {R} prompt sets cannot be split over {n_img} images")
        pe = image_pe.float().contiguous().reshape(-1)
        if pe.numel() != 256 * 64 * 64:
            raise _lib.UllsamError("decoder: image_pe must be [1,256,64,64]")
        sp = sparse.float().contiguous()
        Ns = sp.shape[1]
        d, mode, dbatch = self._dense_args(dense, R)
        if dbatch not in (1, R):
            raise _lib.UllsamError(f"decoder: dense prompt batch {dbatch} must be 1 or {R}")
        if multimask_output is None:
            masks = torch.empty((R, 4, 256, 256), dtype=torch.float32, device=self.device)
            iou = torch.empty((R, 4), dtype=torch.float32, device=self.device)
            rc = self.lib.ullsam_decoder_predict_masks(self._h, ptr(emb), n_img, ptr(pe), ptr(sp), P, Ns, ptr(d), mode,
                                                       dbatch, ptr(masks), ptr(iou), self._stream())
        else:
            cnt = 3 if multimask_output else 1
            masks = torch.empty((R, cnt, 256, 256), dtype=torch.float32, device=self.device)
            iou = torch.empty((R, cnt), dtype=torch.float32, device=self.device)
            rc = self.lib.ullsam_decoder_forward(self._h, ptr(emb), n_img, ptr(pe), ptr(sp), P, Ns, ptr(d), mode, dbatch,
                                                 1 if multimask_output else 0, ptr(masks), ptr(iou), self._stream())
        check(rc, self._h, "ullsam_decoder_forward")
        return masks, iou

    def transformer_forward(self, image_embedding: torch.Tensor, image_pe: torch.Tensor, point_embedding: torch.Tensor
                            ) -> Tuple[torch.Tensor, torch.Tensor]:
        """TwoWayTransformer.forward: ([R,256,64,64], [R|1,256,64,64], [R,Nt,256]) -> (queries, keys)."""
        src = image_embedding.float().contiguous()
        R = src.shape[0]
        pe = image_pe[:1].float().contiguous()
        pt = point_embedding.float().contiguous()
        if pt.shape[0] != R:
            pt = pt.expand(R, -1, -1).contiguous()
        Nt = pt.shape[1]
        q = torch.empty((R, Nt, 256), dtype=torch.float32, device=self.device)
        k = torch.empty((R, 4096, 256), dtype=torch.float32, device=self.device)
        check(self.lib.ullsam_transformer_forward(self._h, ptr(src), R, ptr(pe), ptr(pt), Nt, ptr(q), ptr(k),
                                                  self._stream()), self._h, "ullsam_transformer_forward")
        return q, k

    def segment_host(self, images: torch.Tensor, image_pe: torch.Tensor, sparse: torch.Tensor, dense_vec: torch.Tensor,
                     multimask_output: bool, masks_out: torch.Tensor, iou_out: torch.Tensor,
                     emb_out: Optional[torch.Tensor] = None) -> None:
        """End to end on HOST tensors (pinned recommended): H2D, encoder, decoder (one prompt set per image), D2H."""
        for t in (images, image_pe, sparse, dense_vec, masks_out, iou_out):
            assert not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        B, Ns = images.shape[0], sparse.shape[1]
        check(self.lib.ullsam_segment_host(self._h, ptr(images), B, ptr(image_pe), ptr(sparse), Ns, ptr(dense_vec),
                                           1 if multimask_output else 0, ptr(masks_out), ptr(iou_out), ptr(emb_out)),
              self._h, "ullsam_segment_host")
