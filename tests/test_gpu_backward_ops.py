"""Backward building blocks (weight / data gradients) through the C ABI against torch autograd in fp32."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import gpu_ops  # noqa: E402


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


WGRAD_SHAPES = [
    # B, H, W, Cin, Cout, k, stride, dil
    (300, 1, 1, 256, 256, 1, 1, 1),       # linear (rows as images), one K split tile set
    (1111, 1, 1, 256, 2048, 1, 1, 1),     # FFN linear1, ragged row count
    (700, 1, 1, 2048, 256, 1, 1, 1),      # FFN linear2
    (5, 62, 8, 512, 128, 1, 1, 1),        # layer2 conv1
    (5, 62, 8, 128, 128, 3, 1, 1),        # layer2 conv2 (3x3)
    (3, 124, 16, 128, 128, 3, 2, 1),      # layer2.0 conv2 (3x3 stride 2)
    (3, 124, 16, 256, 512, 1, 2, 1),      # layer2.0 downsample (1x1 stride 2)
    (4, 31, 4, 512, 512, 3, 1, 2),        # layer4 conv2, dilation 2
    (3, 63, 8, 256, 256, 3, 2, 1),        # layer3.0 conv2, odd height
    (2, 31, 4, 64, 128, 1, 1, 1),         # 64 input channels (BLOCK_N = 64 variant)
]


@pytest.mark.parametrize("shape", WGRAD_SHAPES)
def test_conv_wgrad_tcgen05(shape):
    B, H, W, Cin, Cout, k, stride, dil = shape
    pad = dil if k == 3 else 0
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, Cin, H, W, generator=g).bfloat16().float()
    Ho, Wo = gpu_ops.conv_out(H, k, stride, pad, dil), gpu_ops.conv_out(W, k, stride, pad, dil)
    dy = torch.randn(B, Cout, Ho, Wo, generator=g).bfloat16().float()
    ref = torch.nn.grad.conv2d_weight(x.cuda(), (Cout, Cin, k, k), dy.cuda(), stride=stride, padding=pad, dilation=dil)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    dyd = dy.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    dw = gpu_ops.conv_wgrad(xd, dyd, k, stride, dil, pad)               # [Cout, k, k, Cin]
    torch.cuda.synchronize()
    assert rel_err(dw.permute(0, 3, 1, 2), ref) < 2e-5
