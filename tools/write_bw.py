#!/usr/bin/env python
"""HBM write / copy bandwidth probes with plain torch fills (context for the store-bound GEMM layers)."""
import torch
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for mb in (32, 130, 260, 1024):
    x = torch.empty(mb * 1000 * 1000, dtype=torch.uint8, device="cuda")
    y = torch.empty_like(x)
    ms = t(lambda: x.zero_())
    mc = t(lambda: y.copy_(x))
    print(f"{mb} MB  fill {ms*1e3:.1f} us = {mb/ms:.0f} GB/s   copy {mc*1e3:.1f} us = {2*mb/mc:.0f} GB/s (r+w)")
