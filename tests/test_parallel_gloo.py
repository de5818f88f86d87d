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
