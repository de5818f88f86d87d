"""Fused layer1 bottleneck tail (csrc/bneck_fused.cu: conv2 3x3 64->64 + ReLU -> conv3 1x1 64->256 + identity + ReLU in one launch)
against torch on identical bf16 operands with the kernel's rounding points (h2 and the output in bf16), through the C ABI
(sedt_op_bneck_tail), and against the two separate launches it replaces.  GPU only."""
import pytest
import torch
import torch.nn.functional as F

import gpu_ops
from sound_event_detection_transformer_b200 import _lib

pytestmark = pytest.mark.gpu


def r16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def make(B, H, W, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g, device="cuda") * std
    h1 = F.relu(rn(B, H, W, 64)).to(torch.bfloat16)                     # NHWC, post-ReLU like the real conv1 output
    w2 = rn(64, 3, 3, 64, std=(2.0 / 576) ** 0.5).to(torch.bfloat16)    # [Cout][R][S][Cin]
    w3 = rn(256, 64, std=(2.0 / 64) ** 0.5 * 0.35).to(torch.bfloat16)
    b2, b3 = rn(64, std=0.1), rn(256, std=0.1)
    res = F.relu(rn(B, H, W, 256)).to(torch.bfloat16)
    return h1, w2, b2, w3, b3, res


def reference(h1, w2, b2, w3, b3, res):
    x = h1.float().permute(0, 3, 1, 2)
    h2 = r16(F.relu(F.conv2d(x, w2.float().permute(0, 3, 1, 2), b2, padding=1)))
    y = F.conv2d(h2, w3.float().view(256, 64, 1, 1), b3) + res.float().permute(0, 3, 1, 2)
    return r16(F.relu(y)).permute(0, 2, 3, 1).contiguous()


def run(h1, w2, b2, w3, b3, res):
    lib = _lib.load()
    B, H, W, _ = h1.shape
    out = torch.full((B, H, W, 256), float("nan"), dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.sedt_op_bneck_tail(h1.data_ptr(), w2.data_ptr(), b2.data_ptr(), w3.data_ptr(), b3.data_ptr(), res.data_ptr(),
                                      out.data_ptr(), B, H, W, _lib.current_stream()))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("B,H,W", [(3, 125, 16), (2, 32, 16), (1, 124, 16), (5, 16, 8), (40, 124, 16), (7, 61, 16)])
def test_bneck_tail_matches_torch(B, H, W):
    args = make(B, H, W, 10 * B + H)
    got = run(*args).float()
    want = reference(*args)
    assert torch.isfinite(got).all()
    err = ((got - want).norm() / want.norm()).item()
    print(f"bneck_tail B={B} H={H} W={W}: rel-L2 {err:.2e}, max abs {(got - want).abs().max().item():.2e}")
    assert err < 3e-3                               # bf16 output rounding (ulp flips from the fp32 summation order)


def test_bneck_tail_matches_the_two_launches_and_is_deterministic():
    h1, w2, b2, w3, b3, res = make(9, 124, 16, 77)
    a = run(h1, w2, b2, w3, b3, res)
    assert torch.equal(a, run(h1, w2, b2, w3, b3, res))
    h2 = gpu_ops.conv(h1, w2, bias=b2, stride=1, dil=1, pad=1, relu=True)
    two = gpu_ops.conv(h2, w3.view(256, 1, 1, 64), bias=b3, residual=res, relu=True)
    torch.cuda.synchronize()
    err = ((a.float() - two.float()).norm() / two.float().norm()).item()
    assert err < 2e-3, err
