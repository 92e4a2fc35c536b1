"""Keeps an Engine's packed device weights in step with an nn.Module's parameters."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib
from ..engine import Engine


class EngineBound:
    """Mixin: lazily creates the Engine on first CUDA forward and re-uploads tensors whose storage or version
    changed (load_state_dict, .to(), optimizer steps made elsewhere...)."""

    _engine: Optional[Engine] = None
    _uploaded: Dict[str, Tuple[int, int, torch.dtype]]

    def _engine_config(self):  # -> ctypes Config
        raise NotImplementedError

    def _key_prefix(self) -> str:
        raise NotImplementedError

    def _get_engine(self, device: torch.device) -> Engine:
        if device.type != "cuda":
            raise _lib.UllsamError(
                f"{type(self).__name__}: input is on {device}; this path only runs as sm_100a CUDA kernels (no CPU fallback)")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        eng = self.__dict__.get("_engine")
        if eng is None or eng.device_index != idx:
            if eng is not None:
                eng.close()
            eng = Engine(config=self._engine_config(), device=idx)
            self.__dict__["_engine"] = eng
            self.__dict__["_uploaded"] = {}
        self._sync_weights(eng)
        return eng

    def _sync_weights(self, eng: Engine) -> None:
        up = self.__dict__["_uploaded"]
        prefix = self._key_prefix()
        for name, t in self.state_dict(keep_vars=True).items():
            sig = (t.data_ptr(), t._version, t.dtype)
            if up.get(name) == sig:
                continue
            eng.set_weight(prefix + name, t)
            up[name] = sig

    def _check_no_grad(self, *tensors) -> None:
        """The kernels return tensors with no autograd graph.  Raise whenever a backward could be expected from this
        call: grad mode on and either an input or one of THIS module's parameters requires grad — the reference's
        training loop (train_joint_v2.py:1015-1063) feeds inputs that never require grad into modules whose parameters
        do, so checking the inputs alone would silently stop the SAM branch from training."""
        if not torch.is_grad_enabled():
            return
        if any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors) or \
                any(p.requires_grad for p in self.parameters()):
            raise _lib.UllsamError(
                f"{type(self).__name__} (sm_100a) implements the forward pass only (BASELINE north_star): grad mode is "
                "on and an input or a parameter of this module requires grad, so a backward would be expected and "
                "there is none (no autograd fallback).  Call it under torch.no_grad() / torch.inference_mode(), or "
                "freeze the module with .requires_grad_(False) and detach the inputs")
