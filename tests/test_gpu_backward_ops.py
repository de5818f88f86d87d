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


DGRAD_SHAPES = [
    # B, H, W, Cin, Cout, k, stride, dil, mask
    (200, 1, 1, 256, 2048, 1, 1, 1, True),     # FFN linear1 data gradient with the ReLU mask of ... (any mask)
    (300, 1, 1, 2048, 256, 1, 1, 1, False),    # FFN linear2
    (4, 62, 8, 128, 128, 3, 1, 1, True),       # layer2 conv2
    (3, 124, 16, 128, 128, 3, 2, 1, True),     # layer2.0 conv2, stride 2 (zero insertion)
    (3, 124, 16, 256, 512, 1, 2, 1, False),    # layer2.0 downsample, 1x1 stride 2
    (3, 31, 4, 512, 512, 3, 1, 2, True),       # layer4 conv2, dilation 2
    (2, 63, 8, 256, 256, 3, 2, 1, False),      # layer3.0 conv2 with odd height
]


@pytest.mark.parametrize("shape", DGRAD_SHAPES)
def test_conv_dgrad_through_forward_kernels(shape):
    B, H, W, Cin, Cout, k, stride, dil, use_mask = shape
    pad = dil if k == 3 else 0
    g = torch.Generator().manual_seed(9)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5)
    scale = torch.rand(Cout, generator=g) + 0.5
    Ho, Wo = gpu_ops.conv_out(H, k, stride, pad, dil), gpu_ops.conv_out(W, k, stride, pad, dil)
    dy = torch.randn(B, Cout, Ho, Wo, generator=g).bfloat16().float()
    act = torch.randn(B, Cin, H, W, generator=g).bfloat16().float()
    weff = (w * scale[:, None, None, None]).bfloat16().float()
    ref = torch.nn.grad.conv2d_input((B, Cin, H, W), weff.cuda(), dy.cuda(), stride=stride, padding=pad, dilation=dil)
    if use_mask:
        ref = ref * (act.cuda() > 0)
    dyd = dy.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    maskd = act.permute(0, 2, 3, 1).contiguous().cuda().bfloat16() if use_mask else None
    out = gpu_ops.conv_dgrad(dyd, w, (H, W), scale=scale.cuda(), stride=stride, dil=dil, mask=maskd)
    torch.cuda.synchronize()
    assert rel_err(out.permute(0, 3, 1, 2), ref) < 4e-3          # bf16 output rounding


def test_relu_mask_and_colsum():
    g = torch.Generator().manual_seed(1)
    act = torch.randn(1000, 256, generator=g).cuda().bfloat16()
    g1 = torch.randn(1000, 256, generator=g).cuda().bfloat16()
    g2 = torch.randn(1000, 256, generator=g).cuda().bfloat16()
    out = gpu_ops.relu_mask(act, g1, g2)
    ref = ((g1.float() + g2.float()) * (act.float() > 0)).bfloat16()
    assert torch.equal(out, ref)
    out1 = gpu_ops.relu_mask(act, g1)
    assert torch.equal(out1, (g1.float() * (act.float() > 0)).bfloat16())
    for t in (g1, g1.float()):
        s = gpu_ops.colsum(t)
        assert rel_err(s, t.double().sum(0)) < 1e-5
    wide = torch.randn(7, 21 * 256, generator=g).cuda().bfloat16()          # batch sum of per-query rows (query_embed gradient)
    assert rel_err(gpu_ops.colsum(wide), wide.double().sum(0)) < 1e-5


@pytest.mark.parametrize("rows", [1, 37, 5000])
def test_layernorm_backward(rows):
    g = torch.Generator().manual_seed(rows)
    x = (torch.randn(rows, 256, generator=g) * 2 + 0.3).cuda().requires_grad_(True)
    gamma = (torch.rand(256, generator=g) + 0.5).cuda().requires_grad_(True)
    beta = torch.randn(256, generator=g).cuda().requires_grad_(True)
    g1 = torch.randn(rows, 256, generator=g).cuda().bfloat16()
    g2 = torch.randn(rows, 256, generator=g).cuda().bfloat16()
    g3 = torch.randn(rows, 256, generator=g).cuda()
    dres = torch.randn(rows, 256, generator=g).cuda()
    y = torch.nn.functional.layer_norm(x, (256,), gamma, beta, 1e-5)
    y.backward(g1.float() + g2.float() + g3)
    dx, dg, db = gpu_ops.layernorm_bwd(x.detach(), gamma.detach(), g1, g2, g3, dres)
    torch.cuda.synchronize()
    assert rel_err(dx, x.grad + dres) < 1e-5
    assert rel_err(dg, gamma.grad) < 1e-5 and rel_err(db, beta.grad) < 1e-5
    dx1, _, _ = gpu_ops.layernorm_bwd(x.detach(), gamma.detach(), g1)
    x.grad = None
    torch.nn.functional.layer_norm(x, (256,), gamma, beta, 1e-5).backward(g1.float())
    assert rel_err(dx1, x.grad) < 1e-5


@pytest.mark.parametrize("engine", [0, 1])
@pytest.mark.parametrize("case", [(5, 124, 124, False, False), (4, 21, 124, True, False), (3, 21, 21, False, True),
                                  (2, 128, 128, True, False), (3, 11, 128, False, False), (2, 70, 50, True, True)])
def test_attention_backward(case, engine):
    B, Lq, Lk, use_kpm, use_amask = case
    E, nh = 256, 8
    g = torch.Generator().manual_seed(Lq * 7 + Lk)
    q = torch.randn(B, Lq, E, generator=g).cuda().bfloat16()
    k = torch.randn(B, Lk, E, generator=g).cuda().bfloat16()
    v = torch.randn(B, Lk, E, generator=g).cuda().bfloat16()
    do = torch.randn(B, Lq, E, generator=g).cuda().bfloat16()
    kpm = None
    if use_kpm:
        kpm = torch.zeros(B, Lk, dtype=torch.bool)
        for b in range(B):
            kpm[b, Lk - 1 - 3 * b:] = True
        kpm = kpm.cuda()
    amask = None
    if use_amask:
        amask = torch.zeros(Lq, Lk)
        amask[: Lq // 2, Lk // 2:] = float("-inf")
        amask[Lq // 2:, : Lk // 2] = float("-inf")
        amask = amask.cuda()
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    def heads(t, L):
        return t.view(B, L, nh, E // nh).transpose(1, 2)
    s = heads(qf, Lq) @ heads(kf, Lk).transpose(-1, -2) * (E // nh) ** -0.5
    if amask is not None:
        s = s + amask
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
    o = (torch.softmax(s, -1) @ heads(vf, Lk)).transpose(1, 2).reshape(B, Lq, E)
    o.backward(do.float())
    dq, dk, dv = gpu_ops.attention_bwd(q, k, v, do, nh, kpm, amask, engine=engine)
    torch.cuda.synchronize()
    assert rel_err(dq, qf.grad) < 6e-3 and rel_err(dk, kf.grad) < 6e-3 and rel_err(dv, vf.grad) < 6e-3
