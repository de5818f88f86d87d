#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs on one GPU (the headline config 2 is bench.py):
  c1  SEDT E=3, Q=10, T=500, batch 64 eval forward (bf16 and fp32 tiers)
  c3  HungarianMatcher, 8192 clips, Q=20, K~U{0..10}: cost blocks + assignment in one launch
  c5  SP-SEDT pretraining forward, batch 200 + 10 patches of 128x64 per clip
Prints one JSON line per config.  CUDA-event timed, warm-up 3, inputs resident on the device."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from sound_event_detection_transformer_b200 import _lib, flops, spec, synth  # noqa: E402
from sound_event_detection_transformer_b200.sedt import build_matcher, build_model  # noqa: E402


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def model_for(cfg, precision, seed):
    args = spec.config_args(cfg)
    args.precision = precision
    model, _, _ = build_model(args)
    model.load_state_dict(synth.synth_state_dict(args, seed), strict=True)
    return args, model.cuda().eval()


def main():
    which = sys.argv[1:] or ["c1", "c3", "c5"]
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    with torch.no_grad():
        if "c1" in which:
            for precision in ("bf16", "fp32"):
                args, model = model_for("c1", precision, 11)
                x = synth.synth_clips(64, 500, 64, seed=1).cuda()
                ms = timed(lambda: model(x), 20 if precision == "bf16" else 5)
                fl = flops.forward_flops_per_clip(args, 500)["total"]
                print(json.dumps({"config": "c1 SEDT E=3 Q=10 T=500 B=64 eval forward", "precision": precision, "ms_per_step": ms,
                                  "clips_per_s": 64 / ms * 1e3, "tflops": fl * 64 / ms / 1e9}))
                del model
        if "c3" in which:
            outputs, targets = synth.synth_matcher_case(8192, 20, 10, 0, 10, seed=3)
            o = {k: v.cuda() for k, v in outputs.items()}
            t = [{k: v.cuda() for k, v in tg.items()} for tg in targets]
            matcher = build_matcher(spec.default_args())
            t0 = time.perf_counter(); matcher(o, t); torch.cuda.synchronize(); first = time.perf_counter() - t0
            t0 = time.perf_counter()
            for _ in range(5):
                matcher(o, t)                                   # reference-shaped call: host packing + D2H of indices
            torch.cuda.synchronize()
            api_ms = (time.perf_counter() - t0) / 5 * 1e3
            # the same public call with the targets packed once per batch (HungarianMatcher.pack_targets: what a training loop does,
            # one pack for the matcher calls of all decoder layers) -- indices come back as two [B, Q] matrices with lazy per-clip views
            packed = matcher.pack_targets(t, o["pred_logits"].device)
            matcher(o, packed); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                idx, _coef = matcher(o, packed)
            torch.cuda.synchronize()
            api_packed_ms = (time.perf_counter() - t0) / 20 * 1e3
            matcher.device_indices = True
            t0 = time.perf_counter()
            for _ in range(20):
                matcher(o, packed)
            torch.cuda.synchronize()
            api_packed_dev_ms = (time.perf_counter() - t0) / 20 * 1e3
            matcher.device_indices = False
            lib = _lib.load()
            # kernel only: pre-packed device buffers
            sizes = [len(v["boxes"]) for v in t]
            tgt_ids = torch.cat([v["labels"] for v in t]).contiguous(); tgt_box = torch.cat([v["boxes"] for v in t]).contiguous()
            off = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int32).cuda()
            rows = torch.empty(8192, 20, dtype=torch.int64, device="cuda"); cols = torch.empty_like(rows)
            counts = torch.empty(8192, dtype=torch.int32, device="cuda"); status = torch.zeros(1, dtype=torch.int32, device="cuda")
            lg, bx = o["pred_logits"].contiguous(), o["pred_boxes"].contiguous()

            def kern():
                _lib.check(lib.sedt_matcher(lg.data_ptr(), bx.data_ptr(), tgt_ids.data_ptr(), tgt_box.data_ptr(), off.data_ptr(),
                                            8192, 20, 11, max(sizes), 1.0, 5.0, 2.0, None, max(sizes), rows.data_ptr(),
                                            cols.data_ptr(), counts.data_ptr(), status.data_ptr(), _lib.current_stream()))
            k_ms = timed(kern, 50)
            bytes_per_clip = 20 * 11 * 4 + 20 * 2 * 4 + 5 * 16 + 2 * 20 * 8 + 4
            print(json.dumps({"config": "c3 matcher 8192 clips Q=20 K~U{0..10}", "kernel_ms": k_ms, "kernel_clips_per_s": 8192 / k_ms * 1e3,
                              "kernel_GBps": 8192 * bytes_per_clip / k_ms / 1e6, "hbm_frac": 8192 * bytes_per_clip / k_ms / 1e6 / peak.get("hbm_gbs", 6541.5),
                              "api_ms": api_ms, "api_clips_per_s": 8192 / api_ms * 1e3, "first_call_ms": first * 1e3,
                              "api_packed_targets_ms": api_packed_ms, "api_packed_targets_clips_per_s": 8192 / api_packed_ms * 1e3,
                              "api_packed_targets_device_indices_ms": api_packed_dev_ms}))
        if "c5" in which:
            args, model = model_for("c5", "bf16", 15)
            B = 200
            x = synth.synth_clips(B, 496, 64, seed=9).cuda()
            mask = torch.zeros(B, 496, 64, dtype=torch.bool, device="cuda")
            patches = synth.synth_patches(B, 10, 128, 64, seed=9).cuda()
            from sound_event_detection_transformer_b200.utils import NestedTensor
            nt = NestedTensor(x, mask); nt.unpadded = True
            lib = _lib.load()
            ms = timed(lambda: model(nt, patches), 10)
            import ctypes as C
            lib.sedt_profile_enable(1)
            for _ in range(3):
                model(nt, patches)
            msc = (C.c_double * len(_lib.KERNEL_CLASSES))(); nc = (C.c_longlong * len(_lib.KERNEL_CLASSES))()
            lib.sedt_profile_read(msc, nc); lib.sedt_profile_enable(0)
            print(json.dumps({"config": "c5 SP-SEDT forward B=200, 10 patches 128x64", "ms_per_step": ms, "clips_per_s": B / ms * 1e3,
                              "tflops": 30.67e9 * B / ms / 1e9,
                              "per_class_ms": {n: msc[i] / 3 for i, n in enumerate(_lib.KERNEL_CLASSES) if nc[i]}}))


if __name__ == "__main__":
    main()
