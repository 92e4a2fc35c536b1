"""Multi-process plumbing on CPU: world_size-2 gloo run of the batch sharding + all-gather used at N > 1 GPUs."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_the_batch():
    from ullsam_b200.parallel import shard_range, shard_sizes

    for n in (0, 1, 7, 32, 33, 256):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(shard_sizes(n, w)) - min(shard_sizes(n, w)) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


class _FakeEngine:
    """Stands in for ullsam_b200.engine.Engine on CPU: deterministic functions of the inputs, so the gathered
    result can be checked against a single-process evaluation."""

    device = torch.device("cpu")

    def encoder_forward(self, images):
        return images.mean(dim=(1, 2, 3), keepdim=True).expand(-1, 4, 2, 2).contiguous()

    def decoder_forward(self, emb, pe, sparse, dense, multimask, prompts_per_image=1):
        if emb.shape[0] == 0:   # like the C ABI: "decoder: n_img, P must be positive"
            raise RuntimeError("decoder: n_img, P must be positive")
        m = emb[:, :1] + sparse.sum(dim=(1, 2)).reshape(-1, 1, 1, 1) + dense.sum() + pe.sum()
        return m.contiguous(), m.mean(dim=(2, 3))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ullsam_b200.parallel import gather_batch, segment_sharded, shard_range

    g = torch.Generator().manual_seed(0)
    images = torch.rand(n, 3, 8, 8, generator=g)
    sparse = torch.rand(n, 2, 4, generator=g)
    pe, dense = torch.rand(1, 4, 2, 2, generator=g), torch.rand(1, 4, generator=g)
    masks, iou, emb = segment_sharded(_FakeEngine(), images, pe, sparse, dense, False, gather_embeddings=True)
    lo, hi = shard_range(n, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1)
    ids = gather_batch(local, n)
    if rank == 0:
        q.put((masks, iou, emb, ids))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 7, 1])   # 1: fewer images than ranks, rank 1 holds an empty shard
def test_two_rank_gloo_shard_and_gather(n):
    from ullsam_b200.parallel import segment_sharded

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    masks, iou, emb, ids = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    images = torch.rand(n, 3, 8, 8, generator=g)
    sparse = torch.rand(n, 2, 4, generator=g)
    pe, dense = torch.rand(1, 4, 2, 2, generator=g), torch.rand(1, 4, generator=g)
    m1, s1, e1 = segment_sharded(_FakeEngine(), images, pe, sparse, dense, False, gather_embeddings=True)
    assert torch.equal(ids.flatten(), torch.arange(n, dtype=torch.float32))
    assert torch.allclose(masks, m1) and torch.allclose(iou, s1) and torch.allclose(emb, e1)
