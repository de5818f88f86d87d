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
    # SEDT_BENCH_NO_ALLREDUCE=1: the same N-rank step without the gradient exchange (names the exposed all-reduce time by difference)
    model.grad_allreduce = world > 1 and not os.environ.get("SEDT_BENCH_NO_ALLREDUCE")
    model.grad_allreduce_overlap = bool(os.environ.get("SEDT_ALLREDUCE_OVERLAP"))
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
    # ---- end to end: clips from pinned host memory through a side-stream prefetcher (data_utils/DataLoad.py:304-336), the step
    # through the public module / criterion / optimizer calls, the scalar loss read back to pinned host memory every step
    e2e = clocks = None
    if getattr(o, "bench_keys", False):
        from sound_event_detection_transformer_b200.prefetch import ClipPrefetcher
        host = [synth.synth_clips(B, 496, 64, seed=300 + 8 * rank + i).pin_memory() for i in range(4)]
        pf = ClipPrefetcher((host[i % 4] for i in range(o.warmup + o.steps + 4)), dev)
        loss_host = torch.zeros(1).pin_memory()

        def e2e_step():
            nonlocal x
            x = pf.next()
            loss_host.copy_(step().detach().reshape(1), non_blocking=True)
        x_keep = x
        for _ in range(o.warmup):
            e2e_step()
        torch.cuda.synchronize(); parallel.barrier()
        sampler = None
        if rank == 0:
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            from bench import ClockSampler
            sampler = ClockSampler(local); sampler.start()
        # keep every GPU under the bench load while nvidia-smi starts sampling: the SAME number of extra steps on every rank
        # (a step contains the gradient all-reduce: a rank-dependent count would dead-lock the collective)
        for _ in range(40):
            x = x_keep; step()
        torch.cuda.synchronize()
        if sampler is not None:
            sampler.mark()
        parallel.barrier()
        t0 = time.perf_counter()
        for _ in range(o.steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_ms = parallel.max_over_ranks((time.perf_counter() - t0) * 1e3, dev) / o.steps
        clocks = sampler.stop() if sampler is not None else None
        x = x_keep
        e2e = {"value": world * B / e2e_ms * 1e3, "unit": "clips/s", "h2d_bytes_per_step": B * 496 * 64 * 4, "d2h_bytes_per_step": 4,
               "ms_per_step": e2e_ms}
    if rank != 0:
        return None
    fl = flops.forward_flops_per_clip(args, 496, 64)["total"]
    step_flops = 3 * fl - 0.99e9            # fwd + dgrad + wgrad minus the frozen conv1 + layer1 weight gradients
    extra = {}
    if getattr(o, "bench_keys", False):
        pk_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
        pk = json.load(open(pk_path)) if os.path.exists(pk_path) else {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
        ach = B * step_flops / ms / 1e9
        extra["roofline"] = {"bound": "tensor", "kernel": "whole training step (forward + data-gradient + weight-gradient GEMMs are >99% of "
                             "the FLOPs)", "achieved": ach, "peak": float(pk["bf16_tflops_sustained"]), "unit": "TFLOP/s",
                             "frac": ach / float(pk["bf16_tflops_sustained"]), "frac_vs_burst_peak": ach / float(pk["bf16_tflops"]),
                             "traffic": None, "algorithmic_flops_per_clip": step_flops,
                             "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if os.path.exists(pk_path) else "fallback"}
        extra["e2e"], extra["clocks"] = e2e, clocks
        if not getattr(o, "no_cpu_baseline", False):
            ref = run_reference(argparse.Namespace(batch=16, steps=1, warmup=1, gpus=1), emit_line=False)
            extra["cpu_baseline"] = ref["cpu_baseline"]
    return {**_train_line(world, B, o, ms, args), **extra,
            "cuda_graph": not o.no_graph, "grad_allreduce": bool(model.grad_allreduce),
            "allreduce": "none" if not model.grad_allreduce else
                         ("two buckets inside backward, the non-backbone bucket on a side stream overlapping the backbone backward"
                          if model.grad_allreduce_overlap else "one all-reduce of the flat gradient buffer after the last backward kernel"),
            "optimizer": ("FusedAdamW (2 groups) with clip 0.1 fused: csrc/optim.cu, table builds = %d" % opt.table_builds)
            if fused_opt else "torch AdamW (2 groups) + clip_grad_norm_ 0.1 (stock PyTorch)",
            "loss": float(loss.detach()), "gpu_launches": int(model.runtime().kernel_launches() - kl0),
            "kernel_launches_eager_per_step": launches, "phases_ms": phases,
            "achieved_tflops_per_gpu": B * step_flops / ms / 1e9, "per_class": per_class}


TRAIN_METRIC = "clips/sec SEDT E=6 training step (fwd + matcher x3 + set loss + bwd + allreduce + clip + AdamW)"
TRAIN_WORKLOAD = "SEDT E=6, num_queries=20, dec_at, aux_loss, [B,1,496,64] clips, 0..10 events per clip, dropout 0.1"


def _train_line(world, B, o, ms, args):
    return {"metric": TRAIN_METRIC, "value": world * B / ms * 1e3, "unit": "clips/s", "n_gpus": world, "steps": o.steps,
            "warmup": o.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": TRAIN_WORKLOAD, "clips_per_gpu_per_step": B, "global_batch": B * world, "parallelism": f"dp{world}"}}


def run_reference(o, emit_line=True):
    """The reference's own training step on the host cores (rank 0): reference SEDT module in train() mode, reference
    SetCriterion + HungarianMatcher (scipy), autograd backward, torch clip_grad_norm_ + AdamW (engine.py:55-80), fp32."""
    from oracle import ref_loader
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    args = spec.config_args("c2")
    args.dropout = 0.1
    sd = synth.synth_state_dict(args, 12)
    B = o.batch
    if ref_loader.reference_root() is None:
        return {"impl": "reference", "unavailable": "oracle/_ref missing (run oracle/make_ref.py where /root/reference exists); "
                                                    "the oracle port has no training step"}
    model = ref_loader.build_reference_model(args, sd).train()
    criterion = ref_loader.build_reference_criterion(args)
    import numpy as np
    xs = [synth.synth_clips(B, 496, 64, seed=200 + i) for i in range(2)]
    _, targets = synth.synth_matcher_case(B, args.num_queries, args.num_classes, 0, 10, seed=5)
    for t in targets:
        t["orig_size"] = torch.tensor(10.0)
    targets = np.array(targets, dtype=object)
    named = dict(model.named_parameters())
    groups = [{"params": [p for n, p in named.items() if "backbone" not in n and p.requires_grad]},
              {"params": [p for n, p in named.items() if "backbone" in n and p.requires_grad], "lr": 1e-5}]
    opt = torch.optim.AdamW(groups, lr=1e-4, weight_decay=1e-4)
    wd = criterion.weight_dict

    def step(i):
        out = model(xs[i % 2])
        losses, _ = criterion(out, targets, None, slice(B))
        loss = sum(losses[k] * wd[k] for k in losses if k in wd)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)
        opt.step()
        return float(loss.detach())
    for i in range(o.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(o.steps):
        step(i)
    dt = (time.perf_counter() - t0) / o.steps
    v = B / dt
    smp = (f"{B} clips per step x {o.steps} steps of the same workload: the unmodified reference modules (oracle/_ref) in train() mode, "
           f"reference SetCriterion + scipy matcher, autograd, clip_grad_norm_ + AdamW, fp32 torch eager on {cores} host threads")
    line = {"impl": "reference", **_train_line(max(1, getattr(o, "gpus", 1)), B, o, dt * 1e3, args), "dtype": "f32",
            "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": "reference", "sample": smp},
            "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["value"] = v
    return line


if __name__ == "__main__":
    main()
