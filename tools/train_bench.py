#!/usr/bin/env python
"""Config 4: the SEDT E=6 training step (forward + 3 matcher calls + set loss + backward + gradient all-reduce +
clip + AdamW), batch 64 per GPU, dropout 0.  Prints step time, clips/s and a per-phase breakdown.

    python tools/train_bench.py [--batch 64] [--steps 10] [--profile]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_bench.py
"""
import argparse, os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from sound_event_detection_transformer_b200 import spec, synth, parallel, _lib, flops
from sound_event_detection_transformer_b200.sedt import build_model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--dropout", type=float, default=0.1)
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"])
    o = ap.parse_args()
    res = run(o)
    if res is not None:
        print(json.dumps(res))
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


def run(o):
    """o: namespace with batch, steps, warmup, profile, no_graph.  Returns the result dict on rank 0, None elsewhere."""
    rank, world, local = parallel.env_ranks()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    parallel.init_from_env("nccl", dev)
    args = spec.config_args("c2")
    args.dropout = float(getattr(o, "dropout", 0.1))          # the reference's default (train_sedt.py:92)
    sd = synth.synth_state_dict(args, 12)
    model, criterion, _ = build_model(args)
    model.load_state_dict(sd)
    model = model.to(dev).train()
    criterion = criterion.to(dev)
    model.grad_allreduce = world > 1
    model.use_cuda_graph = not o.no_graph
    B = o.batch
    x = synth.synth_clips(B, 496, 64, seed=200 + rank).to(dev)
    _, targets = synth.synth_matcher_case(B, args.num_queries, args.num_classes, 0, 10, seed=5 + rank)
    for t in targets:
        t["labels"] = t["labels"].to(dev); t["boxes"] = t["boxes"].to(dev)
        t["orig_size"] = torch.tensor(10.0, device=dev)
    import numpy as np
    targets_arr = np.array(targets, dtype=object)
    params = [p for p in model.parameters() if p.requires_grad]
    named = dict(model.named_parameters())
    groups = [{"params": [p for n, p in named.items() if "backbone" not in n and p.requires_grad]},
              {"params": [p for n, p in named.items() if "backbone" in n and p.requires_grad], "lr": 1e-5}]
    fused_opt = getattr(o, "optimizer", "fused") == "fused"
    if fused_opt:
        from sound_event_detection_transformer_b200.optim import FusedAdamW
        opt = FusedAdamW(groups, lr=1e-4, weight_decay=1e-4)             # train_sedt.py:234-240,269-270
    else:
        opt = torch.optim.AdamW(groups, lr=1e-4, weight_decay=1e-4)
    wd = criterion.weight_dict
    lib = _lib.load()

    def step(timers=None):
        def mark(name):
            if timers is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(); timers.append((name, e))
        mark("start")
        out = model(x)
        mark("forward")
        losses, _ = criterion(out, targets_arr, None, slice(B))
        loss = sum(losses[k] * wd[k] for k in losses if k in wd)
        mark("matcher+loss")
        opt.zero_grad(set_to_none=True)
        loss.backward()
        mark("backward(+allreduce)")
        if fused_opt:
            opt.step(max_norm=0.1)                                       # engine.py:76-80 in two launches
        else:
            torch.nn.utils.clip_grad_norm_(params, 0.1)                  # engine.py:76-78
            opt.step()
        mark("clip+adamw")
        return loss

    for _ in range(o.warmup):
        step()
    torch.cuda.synchronize(); parallel.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.sedt_launch_count()
    kl0 = model.runtime().kernel_launches()
    e0.record()
    for _ in range(o.steps):
        loss = step()
    e1.record(); torch.cuda.synchronize(); parallel.barrier()
    ms = parallel.max_over_ranks(e0.elapsed_time(e1), dev) / o.steps
    launches = (lib.sedt_launch_count() - l0) / o.steps
    timers = []
    step(timers); torch.cuda.synchronize()
    phases = {timers[i][0]: round(timers[i - 1][1].elapsed_time(timers[i][1]), 3) for i in range(1, len(timers))}
    per_class = None
    if o.profile:
        ms_cls = (C.c_double * len(_lib.KERNEL_CLASSES))(); n_cls = (C.c_longlong * len(_lib.KERNEL_CLASSES))()
        model.use_cuda_graph = False
        step(); torch.cuda.synchronize()
        lib.sedt_profile_enable(1)
        step(); torch.cuda.synchronize()
        _lib.check(lib.sedt_profile_read(ms_cls, n_cls)); lib.sedt_profile_enable(0)
        per_class = {n: {"ms": round(ms_cls[i], 3), "launches": int(n_cls[i])} for i, n in enumerate(_lib.KERNEL_CLASSES) if n_cls[i]}
    if rank != 0:
        return None
    fl = flops.forward_flops_per_clip(args, 496, 64)["total"]
    step_flops = 3 * fl - 0.99e9            # fwd + dgrad + wgrad minus the frozen conv1 + layer1 weight gradients
    return {"metric": "clips/sec SEDT E=6 training step (fwd + matcher x3 + set loss + bwd + allreduce + clip + AdamW)",
            "value": world * B / ms * 1e3, "unit": "clips/s", "n_gpus": world, "steps": o.steps, "warmup": o.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "SEDT E=6, num_queries=20, dec_at, aux_loss, [B,1,496,64] clips, 0..10 events per clip",
                       "dropout": args.dropout,
                       "clips_per_gpu_per_step": B, "cuda_graph": not o.no_graph,
                       "optimizer": ("FusedAdamW (2 groups) with clip 0.1 fused: csrc/optim.cu, table builds = %d" % opt.table_builds)
                       if fused_opt else "torch AdamW (2 groups) + clip_grad_norm_ 0.1 (stock PyTorch)"},
            "loss": float(loss.detach()), "gpu_launches": int(model.runtime().kernel_launches() - kl0),
            "kernel_launches_eager_per_step": launches, "phases_ms": phases,
            "achieved_tflops_per_gpu": B * step_flops / ms / 1e9, "per_class": per_class}


if __name__ == "__main__":
    main()
