"""world_size-2 gloo test of the batch-sharding helpers used by bench.py under torchrun (N > 1)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from sound_event_detection_transformer_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, _ = parallel.init_from_env("gloo")
    lo, hi = parallel.shard_range(n_clips, r, w)
    clips = torch.arange(n_clips, dtype=torch.float32).view(-1, 1) * 2.0       # "work": double every clip id
    local = clips[lo:hi].clone()
    parallel.barrier()
    t_max = parallel.max_over_ranks(10.0 + rank)
    total = parallel.sum_over_ranks(float(hi - lo))
    sizes = [parallel.shard_range(n_clips, i, w)[1] - parallel.shard_range(n_clips, i, w)[0] for i in range(w)]
    parts = parallel.gather_shards(local, sizes)
    if rank == 0:
        q.put((t_max, total, torch.cat(parts).flatten().tolist()))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 7, 1])
def test_two_rank_sharding_gloo(n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    t_max, total, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t_max == 11.0                      # max over ranks, as the bench contract requires
    assert total == n_clips                    # every clip is owned by exactly one rank
    assert gathered == [2.0 * i for i in range(n_clips)]


def test_shard_range_partitions():
    for n in (0, 1, 5, 8, 255, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _grad_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    parallel.init_from_env("gloo")
    # the flat gradient bucket of the training step: each rank holds the gradient of its own clip shard
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    view = flat[10:20].view(2, 5)                       # a parameter's .grad is a view into the bucket
    parallel.allreduce_mean_(flat)
    if rank == 0:
        q.put((flat.tolist(), view.flatten().tolist()))
    torch.distributed.destroy_process_group()


def test_gradient_bucket_allreduce_mean_gloo():
    """The training step's only collective: the bucket is averaged in place, so parameter views see the result."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat, view = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert flat == [1.5 * i for i in range(1000)]
    assert view == [1.5 * i for i in range(10, 20)]


def test_allreduce_mean_single_process_is_identity():
    t = torch.ones(8)
    assert parallel.allreduce_mean_(t) is t and torch.equal(t, torch.ones(8))
