#!/usr/bin/env python
"""Per-parameter gradient error of the native training step vs autograd through the fp32 oracle (debug report)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_train as T
from sound_event_detection_transformer_b200 import spec
args = spec.config_args("c1"); args.enc_layers, args.dec_layers = 2, 2
sd, model, clips, R = T._setup(args, 21, 2, 200)
out = model(clips.cuda())
T._loss(out, {k: v.cuda() for k, v in R.items()}).backward()
named = {n: p for n, p in model.named_parameters() if p.requires_grad}
ref_grads, ref = T._reference_grads(sd, args, clips, R, list(named))
for k in ("pred_logits", "pred_boxes", "at"):
    print(k, ((out[k].detach().cpu() - ref[k]).norm() / ref[k].norm()).item())
for n, p in named.items():
    g, r = p.grad.detach().float().cpu().flatten(), ref_grads[n].flatten()
    rel = ((g - r).norm() / r.norm().clamp_min(1e-20)).item()
    cos = (torch.dot(g, r) / (g.norm() * r.norm()).clamp_min(1e-30)).item()
    print(f"{rel:8.4f} {cos:8.5f} {g.norm().item():10.3e} {r.norm().item():10.3e}  {n}")
