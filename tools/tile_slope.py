#!/usr/bin/env python
"""Per-tile cost of a conv-GEMM shape: time vs batch size -> slope (us per 128-pixel tile per SM) and intercept.
    tile_slope.py H W Cin Cout k [engine]"""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import gpu_ops
H, W, Cin, Cout, k = [int(v) for v in sys.argv[1:6]]
engine = int(sys.argv[6]) if len(sys.argv) > 6 else 1
pad = 1 if k == 3 else 0
pts = []
for B in (64, 128, 256, 512):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, H, W, Cin, generator=g).cuda().bfloat16()
    w = (torch.randn(Cout, k, k, Cin, generator=g) / math.sqrt(Cin * k * k)).cuda().bfloat16()
    bias = torch.randn(Cout, generator=g).cuda()
    for _ in range(3):
        gpu_ops.conv(x, w, None, bias, None, 1, 1, pad, True, torch.bfloat16, engine=engine)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        gpu_ops.conv(x, w, None, bias, None, 1, 1, pad, True, torch.bfloat16, engine=engine)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    tiles = B * H * W / 128 / 148
    pts.append((tiles, us))
    print(f"B={B}: {us:.1f} us, {tiles:.1f} tiles/SM, {2.0*B*H*W*Cout*Cin*k*k/us/1e6:.0f} TFLOP/s")
(t0, u0), (t1, u1) = pts[1], pts[-1]
slope = (u1 - u0) / (t1 - t0)
print(f"slope {slope:.2f} us per tile per SM, intercept {u0 - slope * t0:.1f} us")
