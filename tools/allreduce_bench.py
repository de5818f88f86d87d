#!/usr/bin/env python
"""Time the training step's one exchange in isolation: NCCL sum-all-reduce of the flat fp32 gradient bucket
(36.86 M elements = 147 MB for SEDT E=6) and of its three completion-ordered sub-buckets.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/allreduce_bench.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from sound_event_detection_transformer_b200 import parallel

rank, world, local = parallel.env_ranks()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
parallel.init_from_env("nccl", dev)
res = {}
for name, n in (("all_147MB", 36856576), ("transformer_53MB", 13287296), ("layer4_60MB", 14987264), ("rest_34MB", 8574016)):
    buf = torch.ones(n, device=dev)
    for _ in range(5):
        dist.all_reduce(buf)
    torch.cuda.synchronize(); parallel.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(buf)
    e1.record(); torch.cuda.synchronize()
    ms = parallel.max_over_ranks(e0.elapsed_time(e1) / 20, dev)
    res[name] = {"ms": round(ms, 3), "algbw_GBps": round(n * 4 / ms / 1e6, 1)}
if rank == 0:
    print(json.dumps({"world": world, "allreduce": res}))
dist.destroy_process_group()
