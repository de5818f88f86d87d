#!/usr/bin/env python
"""Benchmark of the SEDT E=6 eval forward hot path (BASELINE.json: configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one forward of a batch of B synthetic log-mel clips [B,1,496,64]
through the SP-SEDT-shaped SEDT E=6 / Q=20 model (random-init weights of the
reference architecture), bf16 operands + fp32 accumulation, per GPU.  Clips
are batch-sharded: every rank runs its own B clips (weak scaling, no
data-path collective).  Rank 0 prints ONE JSON line.

`value`   device-resident inputs, CUDA-event timed, max over ranks.
`e2e`     the public API call model(x) with pinned HOST clips: H2D copy in,
          D2H copy of pred_logits / pred_boxes / at out, inside the timed region.
`roofline` the dominant kernel class (the tcgen05 implicit-GEMM kernel): its
          algorithmic FLOPs per step / its summed launch durations per step,
          measured with CUDA events on the launching stream (a second, profiled
          pass over the same steps); peak = MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the reference's own fp32 PyTorch modules
          (oracle/_ref: an unmodified copy made by oracle/make_ref.py, kind
          "reference"; the oracle port, kind "port", only when that copy is
          missing) on the host cores, the same 256-clip batches per step.
`gpu_eager_baseline`: the same reference modules in stock PyTorch eager on the
          same B200 (fp32, TF32 and bf16 autocast), CUDA-event timed -- the
          comparison SURVEY.md section 2.1 sets; reported, not the target.
`--mode train`: the config-4 training step (tools/train_bench.py) with the same keys.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "clips/sec SEDT E=6 eval fwd at 1/2/4/8 B200 (roofline %) vs ref CPU host path"
UNIT = "clips/s"
T_FRAMES, N_MELS = 496, 64
WORKLOAD = "SP-SEDT-shaped SEDT E=6, num_queries=20, dec_at, DCASE2019-shaped synthetic log-mel [B,1,496,64], eval forward"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc, self.mark_at = gpu_index, [], None, 0

    def mark(self):
        """Samples from here on fall inside the timed region."""
        self.mark_at = len(self.rows)

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        timed = self.rows[self.mark_at:]
        for r in (timed if len(timed) >= 3 else self.rows):     # the GPU is under the same load before the mark
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [v for v in sm if v > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "samples_in_timed_region": len(timed), "interval_ms": 20}


def bench_config(B, world, graph=True, nrot=None, in_bytes=None):
    """`config` of the JSON line: the same dict on both arms (the reference arm runs the same workload)."""
    cfg = {"workload": WORKLOAD, "clips_per_gpu_per_step": B, "global_batch": B * world, "parallelism": f"dp{world}"}
    return cfg


def reference_forward(args, sd, device="cpu", autocast=None):
    """(fn(x) -> outputs, kind): the reference's own SEDT module from oracle/_ref (kind "reference") or, when that copy
    is missing, the oracle port (kind "port", CPU only)."""
    import torch
    from oracle import ref_loader
    if ref_loader.reference_root() is not None:
        model = ref_loader.build_reference_model(args, sd, device)

        def fn(x):
            with torch.no_grad():
                if autocast is not None:
                    with torch.autocast("cuda", dtype=autocast):
                        return model(x)
                return model(x)
        return fn, "reference"
    if str(device) != "cpu":
        return None, "port"
    from oracle import sedt_oracle
    return (lambda x: sedt_oracle.sedt_forward(sd, args, x)), "port"


def cpu_clips_per_sec(args, sd, sample_clips, repeats):
    """The reference forward on the host cores: 1 warm-up + best of `repeats` (cpu_baseline leg of our arm)."""
    import torch
    from sound_event_detection_transformer_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fn, kind = reference_forward(args, sd)
    x = synth.synth_clips(sample_clips, T_FRAMES, N_MELS, seed=77)
    fn(x)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn(x)
        best = min(best, time.perf_counter() - t0)
    return sample_clips / best, cores, best, kind


def gpu_eager_baseline(args, sd, dev, B, steps=5):
    """Stock PyTorch eager of the reference modules on the same GPU (SURVEY 2.1's bar), CUDA-event timed."""
    import torch
    from sound_event_detection_transformer_b200 import synth
    out = {"kind": None, "unit": UNIT, "clips_per_step": B, "steps": steps, "timing": "CUDA events, 2 warm-up steps"}
    x = synth.synth_clips(B, T_FRAMES, N_MELS, seed=78).to(dev)
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for name, allow_tf32, ac in (("fp32", False, None), ("tf32", True, None), ("bf16_autocast", True, torch.bfloat16)):
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = allow_tf32
            fn, kind = reference_forward(args, sd, dev, ac)
            out["kind"] = kind
            if fn is None:
                out["unavailable"] = "oracle/_ref missing: the oracle port is CPU-only"
                return out
            for _ in range(2):
                fn(x)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                fn(x)
            e1.record()
            torch.cuda.synchronize()
            out[name] = B * steps / (e0.elapsed_time(e1) / 1e3)
            del fn
            torch.cuda.empty_cache()
    except Exception as exc:          # a baseline leg must never take the measurement down
        out["error"] = repr(exc)[:200]
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    return out


def run_reference(opts):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores, rank 0 only; the same
    workload, batch and step counts as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from sound_event_detection_transformer_b200 import spec, synth
    if opts.mode == "train":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import train_bench
        return emit(train_bench.run_reference(opts))
    args = spec.config_args("c2")
    sd = synth.synth_state_dict(args, 12)
    B, K, W = opts.batch, opts.steps, opts.warmup
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fn, kind = reference_forward(args, sd)
    xs = [synth.synth_clips(B, T_FRAMES, N_MELS, seed=100 + i) for i in range(2)]
    for i in range(W):
        fn(xs[i % 2])
    t0 = time.perf_counter()
    for i in range(K):
        fn(xs[i % 2])
    dt = time.perf_counter() - t0
    v = B * K / dt
    what = "the unmodified reference modules (oracle/_ref)" if kind == "reference" else "the oracle port of the reference"
    smp = f"{B} clips per step x {K} steps of the same workload, {what}, fp32 torch eager on {cores} host threads"
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": opts.gpus, "steps": K,
        "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(B, max(1, opts.gpus)),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": smp},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


_REAL_STDOUT = None


def _quiet_stdout():
    """Route fd 1 to stderr so that library chatter (e.g. NCCL's version banner) cannot end up next to
    the ONE JSON line; emit() writes that line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="clips per GPU per step (configs[1]: 256)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("SEDT_BENCH_LANES", "2")),
                    help="batches in flight per GPU (eval): lane i is its own model copy + CUDA stream (lanes.py); 1 = plain loop")
    ap.add_argument("--mode", default="eval", choices=["eval", "train"],
                    help="eval: the headline metric (configs[1]); train: the config-4 training step (batch 64 per GPU by default)")
    opts = ap.parse_args()
    opts.warmup = max(opts.warmup, 3)
    if opts.mode == "train" and opts.batch == 256:
        opts.batch = 64                  # config 4: 64 clips per GPU
    if opts.impl == "reference":
        return run_reference(opts)
    if opts.mode == "train":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import train_bench
        import torch.distributed as dist
        opts.profile = False
        opts.dropout = 0.1
        opts.bench_keys = True           # roofline / cpu_baseline / e2e / clocks like the eval line
        res = train_bench.run(opts)
        if res is not None:
            emit(res)
        if dist.is_initialized():
            dist.destroy_process_group()
        return

    import torch
    import torch.distributed as dist
    from sound_event_detection_transformer_b200 import _lib, flops, parallel, spec, synth
    from sound_event_detection_transformer_b200.sedt import build_model

    rank, world, local = parallel.env_ranks()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    parallel.init_from_env("nccl", dev)

    def barrier():
        parallel.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        return parallel.max_over_ranks(ms, dev)

    lib = _lib.load()
    args = spec.config_args("c2")
    args.precision = opts.precision
    sd = synth.synth_state_dict(args, 12)
    model, _, _ = build_model(args)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    model.use_cuda_graph = not opts.no_graph           # the forward's launch sequence is replayed as one CUDA graph
    B, K, W = opts.batch, opts.steps, opts.warmup

    # inputs: rotate over enough distinct batches that they cannot stay L2-resident
    in_bytes = B * T_FRAMES * N_MELS * 4
    nrot = max(2, -(-200 * 2**20 // in_bytes))
    host = [synth.synth_clips(B, T_FRAMES, N_MELS, seed=100 + rank * 16 + i).pin_memory() for i in range(min(nrot, 8))]
    devx = [h.to(dev) for h in host]
    nrot = len(devx)

    from sound_event_detection_transformer_b200.lanes import EvalLanes
    from sound_event_detection_transformer_b200.prefetch import ClipPrefetcher
    nl = max(1, opts.lanes)
    models = [model]
    for _ in range(nl - 1):                              # every lane: its own module, packed weights, workspace and graph
        m, _, _ = build_model(args)
        m.load_state_dict(sd, strict=True)
        m = m.to(dev).eval()
        m.use_cuda_graph = model.use_cuda_graph
        models.append(m)

    def launch_count():
        return int(lib.sedt_launch_count()) + sum(int(m.runtime().graph_kernel_launches) for m in models)

    # ---- (1) device-resident throughput -------------------------------------------------
    def device_throughput(lanes, sample_clocks):
        with torch.no_grad():
            for i in range(max(W, len(lanes))):
                with lanes.stream(i):
                    lanes.model(i)(devx[i % nrot])
            lanes.join()
            barrier()
            sampler = ClockSampler(local) if sample_clocks else None
            if sampler is not None:
                if rank == 0:
                    sampler.start()
                # keep every GPU under the bench load while nvidia-smi starts sampling (untimed extra warm-up, <= 1 s)
                t_w = time.perf_counter()
                while time.perf_counter() - t_w < 1.0 and (time.perf_counter() - t_w < 0.4 or (rank == 0 and len(sampler.rows) < 2)):
                    model(devx[0])
                    torch.cuda.synchronize()
                sampler.mark()
            l0 = launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            lanes.fork()
            for i in range(K):
                with lanes.stream(i):
                    lanes.model(i)(devx[i % nrot])
            lanes.join()
            e1.record()
            barrier()
            t = max_over_ranks(e0.elapsed_time(e1))
            n = launch_count() - l0
            clk = sampler.stop() if (sampler is not None and rank == 0) else None
        return t, n, clk

    # ---- (2) end to end through the public API with host buffers ------------------------
    with torch.no_grad():
        out = model(host[0])
        keys = [k for k in ("pred_logits", "pred_boxes", "at") if k in out]
        d2h = sum(out[k].numel() * out[k].element_size() for k in keys)

    def end_to_end(lanes):
        # the caller-side loop of the reference: batches come from pinned host memory through a side-stream
        # prefetcher (data_utils/DataLoad.py:304-336), the forward runs through the public module call, the
        # results are read back to pinned host memory.  Steady state: the stream of batches is longer than the
        # timed window, so every timed step issues exactly one H2D copy (of a later batch) and one D2H read.
        # Every lane owns its prefetcher and its pinned result buffers; all of a batch's work is ordered on its lane.
        n = len(lanes)
        with torch.no_grad():
            pinned = [{k: torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory() for k in keys} for _ in range(n)]
            pfs = []
            for j in range(n):
                with lanes.stream(j):
                    pfs.append(ClipPrefetcher((host[(j + n * i) % nrot] for i in range((W + K) // n + 6)), dev))

            def e2e_step(i):
                with lanes.stream(i):
                    xb = pfs[i % n].next()
                    o = lanes.model(i)(xb)
                    for k in keys:
                        pinned[i % n][k].copy_(o[k], non_blocking=True)

            for i in range(max(W, n)):
                e2e_step(i)
            lanes.join()
            barrier()
            t0 = time.perf_counter()
            lanes.fork()
            for i in range(K):
                e2e_step(max(W, n) + i)
            lanes.join()
            torch.cuda.synchronize()
            return max_over_ranks((time.perf_counter() - t0) * 1e3)

    lanes = EvalLanes(models, dev)
    ms, launches, clocks = device_throughput(lanes, True)
    value = world * B * K / (ms / 1e3)
    e2e_ms = end_to_end(lanes)
    e2e_value = world * B * K / (e2e_ms / 1e3)
    single = None
    if nl > 1:                                           # the plain one-batch-at-a-time loop, reported next to the headline
        one = EvalLanes([model], dev)
        ms1, _, _ = device_throughput(one, False)
        e2e_ms1 = end_to_end(one)
        single = {"value": world * B * K / (ms1 / 1e3), "ms_per_step": ms1 / K, "e2e_value": world * B * K / (e2e_ms1 / 1e3),
                  "e2e_ms_per_step": e2e_ms1 / K, "unit": UNIT,
                  "note": "one batch in flight (lanes = 1): also the latency of one 256-clip forward"}

    # ---- (3) per-kernel-class durations (profiled pass, same steps) ----------------------
    ms_cls = (C.c_double * len(_lib.KERNEL_CLASSES))()
    n_cls = (C.c_longlong * len(_lib.KERNEL_CLASSES))()
    with torch.no_grad():
        barrier()
        model.use_cuda_graph = False                       # per-kernel events need eager launches
        lib.sedt_profile_enable(1)
        kinds0 = _lib.kernel_kind_counts()
        for i in range(K):
            model(devx[i % nrot])
        _lib.check(lib.sedt_profile_read(ms_cls, n_cls))
        lib.sedt_profile_enable(0)
        kinds1 = _lib.kernel_kind_counts()
    # which of the size-dependent kernels produced these numbers (launches per step of the eager pass = the graph's contents)
    kernel_kinds = {k: (kinds1[k] - kinds0[k]) / K for k in kinds1 if kinds1[k] != kinds0[k]}
    per_class = {n: {"ms_per_step": ms_cls[i] / K, "launches_per_step": n_cls[i] / K}
                 for i, n in enumerate(_lib.KERNEL_CLASSES) if n_cls[i]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fl = flops.forward_flops_per_clip(args, T_FRAMES, N_MELS)
    peaks, peak_src = measured_peaks()
    peak_tf = float(peaks["bf16_tflops_sustained"])
    dom = "gemm_tcgen05" if "gemm_tcgen05" in per_class else max(per_class, key=lambda k: per_class[k]["ms_per_step"])
    dom_ms = per_class[dom]["ms_per_step"]
    dom_flops = fl["tensor_core_gemm"] * B if dom == "gemm_tcgen05" else fl["total"] * B
    achieved = dom_flops / (dom_ms / 1e3) / 1e12
    step_tf = fl["total"] * B * K / (ms / 1e3) / 1e12 / world if world else 0.0
    # DRAM bytes of that kernel class per step: from the committed ncu capture of this same command / batch (ncu cannot run
    # inside the timed region); the newest profiles/r*_class_summary.json, its name and age stated next to the number
    traffic, traffic_src = None, None
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_class_summary.json")))
    if cands and B == 256 and opts.precision == "bf16":
        tp = cands[-1]
        c = json.load(open(tp)).get(dom)
        if c:
            traffic = (c["dram_read_MB"] + c["dram_write_MB"]) * 1e6
            traffic_src = (f"profiles/{os.path.basename(tp)} (ncu dram__bytes_read.sum + dram__bytes_write.sum over the class's "
                           "launches of one step of this command; captured earlier, not in this run)")
    peak_burst = float(peaks.get("bf16_tflops", peak_tf))
    step_ach = fl["total"] * B / ((ms / K) / 1e3) / 1e12
    roofline = {
        "bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": f"{peak_src} bf16_tflops_sustained (the kernel runs inside a long step); burst peak {peak_burst} -> "
                       f"frac {achieved / peak_burst:.3f}",
        "frac_vs_burst_peak": achieved / peak_burst,
        "algorithmic_flops_per_clip": fl["total"], "kernel_flops_per_step": dom_flops, "kernel_ms_per_step": dom_ms,
        "kernel_share_of_step": dom_ms / sum(v["ms_per_step"] for v in per_class.values()),
        "whole_step": {"achieved": step_ach, "frac": step_ach / peak_tf, "frac_vs_burst_peak": step_ach / peak_burst},
        "per_class": per_class,
    }

    cpu = eager = None
    if not opts.no_cpu_baseline:
        v, cores, best, kind = cpu_clips_per_sec(args, sd, 256, 2)
        what = "the unmodified reference modules (oracle/_ref)" if kind == "reference" else "oracle port of the reference"
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"one 256-clip batch of the same workload, {what}, fp32 torch eager on {cores} host threads, "
                         f"1 warm-up + best of 2 ({best:.2f} s per batch)"}
        eager = gpu_eager_baseline(args, sd, dev, B)

    emit({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if opts.precision == "bf16" else "f32", "data": "synthetic",
        "config": bench_config(B, world),
        "cuda_graph": not opts.no_graph,
        "lanes": {"n": nl, "what": "batches in flight per GPU: each lane = its own model copy (weights, workspace, CUDA graph) on its own "
                                   "CUDA stream, batch i on lane i % n; every step is one complete forward of one batch",
                  "single_lane": single},
        "l2": f"inputs rotate over {nrot} distinct {in_bytes / 2**20:.1f} MiB batches; per-step activation traffic (>1 GB) exceeds the 126 MB L2",
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / K},
        "gpu_launches": launches, "kernel_kinds_per_step": kernel_kinds, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "gpu_eager_baseline": eager,
    })
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
