#!/usr/bin/env python
"""Time / profile one conv-GEMM shape through the C ABI:  op_bench.py B H W Cin Cout k [res] [f32] [engine]"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import gpu_ops
B, H, W, Cin, Cout, k = [int(v) for v in sys.argv[1:7]]
res = "res" in sys.argv; f32 = "f32" in sys.argv
engine = 2 if "2sm" in sys.argv else 1
g = torch.Generator().manual_seed(0)
x = torch.randn(B, H, W, Cin, generator=g).cuda().bfloat16()
w = (torch.randn(Cout, k, k, Cin, generator=g) / math.sqrt(Cin * k * k)).cuda().bfloat16()
bias = torch.randn(Cout, generator=g).cuda()
odt = torch.float32 if f32 else torch.bfloat16
r = torch.randn(B, H, W, Cout, generator=g).cuda().to(odt) if res else None
pad = 1 if k == 3 else 0
for _ in range(3):
    gpu_ops.conv(x, w, None, bias, r, 1, 1, pad, True, odt, engine=engine)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    gpu_ops.conv(x, w, None, bias, r, 1, 1, pad, True, odt, engine=engine)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
fl = 2.0 * B * H * W * Cout * Cin * k * k
print(f"{sys.argv[1:]}: {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TFLOP/s")
