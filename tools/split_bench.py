#!/usr/bin/env python
"""Experiment: one CUDA graph of the full batch vs one graph that runs NSPLIT sub-batches on NSPLIT streams."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sound_event_detection_transformer_b200 import spec, synth
from sound_event_detection_transformer_b200.sedt import build_model
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
args = spec.config_args("c2")
sd = synth.synth_state_dict(args, 12)
model, _, _ = build_model(args)
model.load_state_dict(sd); model.cuda().eval()
x = synth.synth_clips(B, 496, 64, seed=3).cuda()
with torch.no_grad():
    model(x)
rt = model.runtime()
def timeit(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def make_graph(nsplit):
    hb = B // nsplit
    parts = []
    for i in range(nsplit):
        need = rt.lib.sedt_workspace_bytes(rt.handle, hb, 496, 64, 0, 0)
        parts.append(dict(x=x[i * hb:(i + 1) * hb].contiguous(), ws=torch.empty(need + 256, dtype=torch.uint8, device="cuda"),
                          res=rt._alloc_outputs(hb, 496, 64, 0, x.device)))
    streams = [torch.cuda.Stream() for _ in range(nsplit - 1)]
    def run():
        main = torch.cuda.current_stream()
        for i, p in enumerate(parts):
            if i == 0:
                rt._launch(p["x"], None, None, 0, 0, p["ws"], p["res"])
            else:
                st = streams[i - 1]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    rt._launch(p["x"], None, None, 0, 0, p["ws"], p["res"])
        for st in streams:
            main.wait_stream(st)
    run(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    return g, parts
for ns in (1, 2, 4):
    g, parts = make_graph(ns)
    ms = timeit(g.replay)
    print(f"nsplit {ns}: {ms:.3f} ms  {B / ms * 1e3:.0f} clips/s")
