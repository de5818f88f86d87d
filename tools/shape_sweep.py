#!/usr/bin/env python
"""Times one conv-GEMM problem (fixed M x K x N) under several NHWC spatial factorizations through the C ABI, to separate
the cost of the GEMM from the cost of the TMA box shape the spatial layout implies.

    shape_sweep.py            # M = 8192-pixel problems (layer4 WITHOUT dilation, e.g. config 1) in several layouts each
    shape_sweep.py --cold     # the same kernels with activations and / or weights rotating out of L2

Findings (B200, round 2): the layout does not matter (box {64,2,16,4} = {64,1,1,128} = {64,4,32,1} to within 2 %), and neither
does L2 residency of the weights or activations (31.2 us for the 3x3 512->512 conv in all four combinations).  Config 2 itself
has dilation = True, so its layer4 runs at 31 x 4 (M = 31744): the 58 / 111 / 80 us launches of profiles/r2o_launches_final.csv
are 1150 / 1350 / 830 TFLOP/s, not slow M = 8192 problems.
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402
import gpu_ops  # noqa: E402

CASES = [
    # name, [(B, H, W)], Cin, Cout, k, residual, out fp32
    ("layer4 conv1 1x1 2048->512", [(256, 16, 2), (8192, 1, 1), (64, 32, 4), (128, 16, 4), (512, 16, 1)], 2048, 512, 1, False, False),
    ("layer4 conv3 1x1 512->2048 +res", [(256, 16, 2), (8192, 1, 1), (64, 32, 4), (512, 16, 1)], 512, 2048, 1, True, False),
    ("layer4 conv2 3x3 512->512", [(256, 16, 2), (64, 32, 4), (128, 16, 4)], 512, 512, 3, False, False),
    ("input_proj 2048->256 f32", [(256, 16, 2), (8192, 1, 1), (64, 32, 4)], 2048, 256, 1, False, True),
    ("layer4.0 conv1 1x1 1024->512 (32x4)", [(256, 32, 4), (32768, 1, 1)], 1024, 512, 1, False, False),
]


def time_case(B, H, W, Cin, Cout, k, res, f32, engine=1, iters=20):
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
    for _ in range(iters):
        gpu_ops.conv(x, w, None, bias, r, 1, 1, pad, True, odt, engine=engine)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms * 1e3, 2.0 * B * H * W * Cout * Cin * k * k / ms / 1e9


def main():
    for name, layouts, Cin, Cout, k, res, f32 in CASES:
        print(name)
        for (B, H, W) in layouts:
            for engine in (1, 2):
                try:
                    us, tf = time_case(B, H, W, Cin, Cout, k, res, f32, engine)
                    print(f"   [{B:>5},{H:>2},{W}] engine {engine}: {us:7.1f} us  {tf:6.0f} TFLOP/s", flush=True)
                except Exception as e:  # noqa: BLE001
                    print(f"   [{B:>5},{H:>2},{W}] engine {engine}: {str(e)[:100]}", flush=True)


def rotate_case(B, H, W, Cin, Cout, k, res, rot_x, rot_w, engine=2, iters=48):
    """Same problem with the activation and / or the weights rotating over enough distinct buffers to defeat the 126 MB L2
    (the state a layer sees inside a forward: activation just written = L2 hits, weights last touched one forward ago = DRAM)."""
    g = torch.Generator().manual_seed(0)
    nx = max(1, int(math.ceil(160e6 / (B * H * W * Cin * 2)))) if rot_x else 1
    nw = max(1, int(math.ceil(160e6 / (Cout * k * k * Cin * 2)))) if rot_w else 1
    xs = [torch.randn(B, H, W, Cin, generator=g).cuda().bfloat16() for _ in range(nx)]
    ws = [(torch.randn(Cout, k, k, Cin, generator=g) / math.sqrt(Cin * k * k)).cuda().bfloat16() for _ in range(nw)]
    bias = torch.randn(Cout, generator=g).cuda()
    r = torch.randn(B, H, W, Cout, generator=g).cuda().bfloat16() if res else None
    pad = 1 if k == 3 else 0
    for i in range(max(nx, nw, 3)):
        gpu_ops.conv(xs[i % nx], ws[i % nw], None, bias, r, 1, 1, pad, True, torch.bfloat16, engine=engine)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        gpu_ops.conv(xs[i % nx], ws[i % nw], None, bias, r, 1, 1, pad, True, torch.bfloat16, engine=engine)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3, nx, nw


def cold():
    for name, (B, H, W), Cin, Cout, k, res in [("layer4 conv1 1x1 2048->512", (256, 16, 2), 2048, 512, 1, False),
                                               ("layer4 conv2 3x3 512->512", (256, 16, 2), 512, 512, 3, False),
                                               ("layer4 conv3 1x1 512->2048 +res", (256, 16, 2), 512, 2048, 1, True),
                                               ("layer3 conv2 3x3 256->256", (256, 31, 4), 256, 256, 3, False),
                                               ("layer3 conv1 1x1 1024->256", (256, 31, 4), 1024, 256, 1, False)]:
        print(name)
        for rot_x, rot_w in ((False, False), (False, True), (True, False), (True, True)):
            for engine in (1, 2):
                try:
                    us, nx, nw = rotate_case(B, H, W, Cin, Cout, k, res, rot_x, rot_w, engine)
                    print(f"   x over {nx:>3} buffers, w over {nw:>3} buffers, engine {engine}: {us:7.1f} us", flush=True)
                except Exception as e:  # noqa: BLE001
                    print(f"   engine {engine}: {str(e)[:100]}", flush=True)


if __name__ == "__main__":
    cold() if "--cold" in sys.argv else main()
