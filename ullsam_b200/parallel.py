"""Data-parallel plumbing for the SAM path: one process per GPU, weights replicated, image batch sharded.

The path has no cross-image operation, so ranks never exchange anything while computing; the only
collective is the final all-gather of mask logits / IoU scores (and, optionally, embeddings).  ``torch.distributed``
is used as-is: backend "nccl" on the GPU box (NVLink 5 / NVSwitch), "gloo" in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``n_items`` owned by ``rank``; the first ``n_items % world`` ranks get one more."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def gather_batch(local: Optional[torch.Tensor], n_items: int, group: Optional[dist.ProcessGroup] = None,
                 device: Optional[torch.device] = None) -> torch.Tensor:
    """All-gather per-rank shards (dim 0 sized by ``shard_range``) back into the full batch, in rank order.
    Equal shards use one ``all_gather_into_tensor``; ragged shards are padded to the largest one first.

    A rank whose shard is empty (``n_items < world``) may pass ``local=None``: it has run no kernel and does not know
    the item shape, so the ranks first agree on (trailing shape, dtype) with a small object all-gather — every rank
    takes that branch together because ``shard_sizes`` is a pure function of (n_items, world)."""
    if not (dist.is_available() and dist.is_initialized()):
        if local is None:
            raise ValueError("gather_batch: an empty shard needs a process group to learn the item shape from")
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_items, world)
    if min(sizes) == 0:
        meta = [None] * world
        dist.all_gather_object(meta, None if local is None else (tuple(local.shape[1:]), local.dtype), group=group)
        known = next((m for m in meta if m is not None), None)
        if known is None:
            raise ValueError("gather_batch: every shard is empty")
        if local is None:
            dev = device if device is not None else torch.device("cpu")
            local = torch.empty((0,) + known[0], dtype=known[1], device=dev)
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError(f"rank {dist.get_rank(group)} holds {local.shape[0]} items, expected {sizes[dist.get_rank(group)]}")
    mx = max(sizes)
    if min(sizes) == mx:
        out = local.new_empty((world * mx,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    padded = local.new_zeros((mx,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    out = local.new_empty((world * mx,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * mx: r * mx + s] for r, s in enumerate(sizes)], dim=0)


class AsyncGather:
    """The path's only collective — the all-gather of a step's mask logits / IoU scores — issued on a
    SIDE stream, so that it overlaps the next step's encoder instead of stalling the compute stream on the slowest rank.

        ag = AsyncGather(device)
        for step in ...:
            masks, iou = decode(...)                      # on the current (compute) stream
            ag.submit([masks, iou], [all_masks[step & 1], all_iou[step & 1]])
        ag.wait()                                         # current stream waits for every gather queued so far

    ``comm_ms()`` returns the device time spent inside the gathers (CUDA events on the side stream) and resets it."""

    def __init__(self, device: torch.device, group: Optional[dist.ProcessGroup] = None, timing: bool = True):
        self.device, self.group, self.timing = device, group, timing
        self.stream = torch.cuda.Stream(device)
        self._spans: List[Tuple[torch.cuda.Event, torch.cuda.Event]] = []

    def submit(self, shards: Sequence[torch.Tensor], outs: Sequence[torch.Tensor]) -> None:
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            if self.timing:
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record(self.stream)
            for shard, out in zip(shards, outs):
                dist.all_gather_into_tensor(out, shard, group=self.group)
                shard.record_stream(self.stream)      # the allocator must not recycle the shard before the gather ran
            if self.timing:
                t1.record(self.stream)
                self._spans.append((t0, t1))

    def wait(self) -> None:
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def comm_ms(self) -> float:
        self.stream.synchronize()
        total = sum(a.elapsed_time(b) for a, b in self._spans)
        self._spans.clear()
        return total


def segment_sharded(engine, images: torch.Tensor, image_pe: torch.Tensor, sparse: torch.Tensor, dense: torch.Tensor,
                    multimask_output: bool, group: Optional[dist.ProcessGroup] = None, gather_embeddings: bool = False):
    """Every rank passes the SAME global batch description (``images`` may be any tensor whose dim 0 is the global
    batch, typically a host tensor); each rank encodes + decodes only its slice and the results are all-gathered.
    Returns (masks, iou[, embeddings]) for the whole batch on every rank."""
    n = images.shape[0]
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_range(n, rank, world)
    dev = engine.device
    if hi > lo:
        emb = engine.encoder_forward(images[lo:hi].to(dev, non_blocking=True))
        d = dense if dense.shape[0] == 1 else dense[lo:hi]
        masks, iou = engine.decoder_forward(emb, image_pe.to(dev), sparse[lo:hi].to(dev), d.to(dev), multimask_output,
                                            prompts_per_image=1)
    else:
        # fewer images than ranks: this rank has nothing to encode (the decoder rejects an empty batch), but it still
        # has to join every collective the other ranks enter, or they hang
        emb = masks = iou = None
    out = (gather_batch(masks, n, group, dev), gather_batch(iou, n, group, dev))
    if gather_embeddings:
        out = out + (gather_batch(emb, n, group, dev),)
    return out
