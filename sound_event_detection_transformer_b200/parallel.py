"""One process per GPU, clips batch-sharded across ranks.

The eval forward, the SP-SEDT forward and the matcher need no data-path
collective: every rank owns a contiguous shard of the clips and a full copy of
the weights.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used
for the rendezvous, barriers and the max-over-ranks reduction of timings /
the gather of per-shard results only.  Mirrors the env-based init of the
reference (utilities/distribute.py:43-65: RANK / WORLD_SIZE / LOCAL_RANK).
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def env_ranks() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_from_env(backend: str = "nccl", device: torch.device | None = None) -> Tuple[int, int, int]:
    rank, world, local = env_ranks()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n clips for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def barrier() -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_shards(local: torch.Tensor, sizes: Sequence[int]) -> List[torch.Tensor] | None:
    """Gather per-rank result shards (dim 0 sizes `sizes`) on rank 0 in clip order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    world, rank = dist.get_world_size(), dist.get_rank()
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    return [o[:s] for o, s in zip(out, sizes)] if rank == 0 else None


def allreduce_mean_(flat: torch.Tensor) -> torch.Tensor:
    """The training step's one exchange: sum-all-reduce of the flat gradient bucket written by
    sedt_backward, divided by the world size (what DistributedDataParallel does bucket by bucket,
    train_spsedt.py:157-158).  In place; a no-op for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
    return flat
