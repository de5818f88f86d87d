#!/usr/bin/env python
"""Fused FFN kernel vs the two-GEMM path on the encoder shape (M = 31744, ff = 2048).  python tools/ffn_bench.py [M]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import gpu_ops
from sound_event_detection_transformer_b200 import _lib
lib = _lib.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 31744
ff = 2048
g = torch.Generator().manual_seed(0)
x = torch.randn(M, 256, generator=g).cuda().bfloat16()
w1 = (torch.randn(ff, 256, generator=g) / 16).cuda().bfloat16(); w2 = (torch.randn(256, ff, generator=g) / 45).cuda().bfloat16()
b1, b2 = torch.randn(ff).cuda(), torch.randn(256).cuda()
res = torch.randn(M, 256, generator=g).cuda(); out = torch.empty_like(res)
def fused():
    _lib.check(lib.sedt_op_ffn(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), res.data_ptr(), out.data_ptr(), M, ff, _lib.current_stream()))
def two():
    h = gpu_ops.conv(x.view(M, 1, 1, 256), w1.view(ff, 1, 1, 256), None, b1, None, 1, 1, 0, True, torch.bfloat16, engine=1)
    gpu_ops.conv(h, w2.view(256, 1, 1, ff), None, b2, res.view(M, 1, 1, 256), 1, 1, 0, False, torch.float32, engine=1)
for name, fn in (("fused", fused), ("two GEMMs", two)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"{name}: {us:.1f} us  {4.0 * M * 256 * ff / us / 1e6:.0f} TFLOP/s")
