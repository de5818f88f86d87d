"""Kernel-level parity: every exported operator (called through the C ABI) against the oracle /
the torch fp32 op it replaces, on seeded inputs.  GPU only."""
import math

import pytest
import torch
import torch.nn.functional as F

import gpu_ops
from oracle import sedt_oracle
from sound_event_detection_transformer_b200 import spec, synth

pytestmark = pytest.mark.gpu


def rel_err(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-12)).item()


def _conv_case(B, H, W, Cin, Cout, k, stride, dil, res, relu, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * math.sqrt(2.0 / (Cin * k * k))
    scale = torch.rand(Cout, generator=g) + 0.5
    bias = torch.randn(Cout, generator=g) * 0.1
    pad = dil if k == 3 else 0
    if dtype == torch.bfloat16:          # the kernel sees bf16-rounded operands; so does the reference
        x, w = x.bfloat16().float(), w.bfloat16().float()
    ref = F.conv2d(x, w, stride=stride, padding=pad, dilation=dil) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    r = None
    if res:
        r = torch.randn(ref.shape, generator=g)
        if dtype == torch.bfloat16:
            r = r.bfloat16().float()
        ref = ref + r
    if relu:
        ref = ref.relu()
    return x, w, scale, bias, pad, r, ref


CONV_SHAPES = [
    # B, H, W, Cin, Cout, k, stride, dil, residual, relu
    (2, 31, 16, 64, 64, 1, 1, 1, False, True),
    (2, 31, 16, 64, 64, 3, 1, 1, False, True),
    (1, 33, 16, 64, 256, 1, 1, 1, True, True),
    (2, 30, 16, 128, 128, 3, 2, 1, False, True),      # layer2.0.conv2 style (stride 2)
    (2, 31, 16, 256, 512, 1, 2, 1, False, False),     # downsample 1x1 stride 2, odd H
    (2, 9, 4, 512, 512, 3, 1, 2, False, True),        # layer4 dilation 2
    (3, 8, 4, 256, 1024, 1, 1, 1, True, True),
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv_simt_fp32(shape):
    B, H, W, Cin, Cout, k, stride, dil, res, relu = shape
    x, w, scale, bias, pad, r, ref = _conv_case(*shape, seed=1)
    out = gpu_ops.conv(x.permute(0, 2, 3, 1).contiguous().cuda(), gpu_ops.repack(w, torch.float32), scale.cuda(), bias.cuda(),
                       r.permute(0, 2, 3, 1).contiguous().cuda() if res else None, stride, dil, pad, relu)
    torch.cuda.synchronize()
    assert out.shape == (B, ref.shape[2], ref.shape[3], Cout)
    assert rel_err(out.permute(0, 3, 1, 2), ref) < 2e-6


def test_linear_simt_small_n():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(77, 256, generator=g)
    w = torch.randn(11, 256, generator=g) / 16
    b = torch.randn(11, generator=g)
    out = gpu_ops.conv(x.view(77, 1, 1, 256).cuda(), w.view(11, 1, 1, 256).cuda(), None, b.cuda())
    torch.cuda.synchronize()
    assert rel_err(out.view(77, 11), F.linear(x, w, b)) < 2e-6


@pytest.mark.parametrize("T", [500, 496, 128, 333, 61])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_stem(T, dtype):
    args = spec.config_args("c1")
    sd = synth.synth_state_dict(args, 21)
    x = synth.synth_clips(2, T, 64, seed=T)
    taps = {}
    body = sedt_oracle.BODY
    ref = F.conv2d(x, sd[body + "conv0.weight"], sd[body + "conv0.bias"])
    ref = F.conv2d(ref, sd[body + "conv1.weight"], stride=2, padding=3)
    ref = F.max_pool2d(F.relu(sedt_oracle.frozen_bn(ref, sd, body + "bn1")), 3, 2, 1)
    out = gpu_ops.stem(x, sd, body, dtype)
    torch.cuda.synchronize()
    assert out.shape == (2, ref.shape[2], 16, 64)
    tol = 3e-6 if dtype == torch.float32 else 4e-3
    assert rel_err(out.permute(0, 3, 1, 2), ref) < tol
    # borders are where the conv0-bias fold could go wrong: check them separately
    o = out.permute(0, 3, 1, 2).float().cpu()
    for sl in (slice(0, 1), slice(-1, None)):
        assert rel_err(o[:, :, sl, :], ref[:, :, sl, :]) < tol * 2
        assert rel_err(o[:, :, :, sl], ref[:, :, :, sl]) < tol * 2


@pytest.mark.parametrize("T", [500, 496, 128, 333, 61, 7])
def test_stem_tcgen05(T):
    """tcgen05 stem (im2col in smem + indicator channel for the conv0-bias border term)."""
    args = spec.config_args("c1")
    sd = synth.synth_state_dict(args, 21)
    x = synth.synth_clips(3, T, 64, seed=T)
    body = sedt_oracle.BODY
    ref = F.conv2d(x, sd[body + "conv0.weight"], sd[body + "conv0.bias"])
    ref = F.conv2d(ref, sd[body + "conv1.weight"], stride=2, padding=3)
    ref = F.max_pool2d(F.relu(sedt_oracle.frozen_bn(ref, sd, body + "bn1")), 3, 2, 1)
    out = gpu_ops.stem(x, sd, body, torch.bfloat16, engine=1)
    torch.cuda.synchronize()
    assert out.shape == (3, ref.shape[2], 16, 64)
    o = out.permute(0, 3, 1, 2).float().cpu()
    assert rel_err(o, ref) < 8e-3                    # bf16 operands (x, folded weights) and bf16 output
    for sl in (slice(0, 1), slice(-1, None)):
        assert rel_err(o[:, :, sl, :], ref[:, :, sl, :]) < 1.6e-2
        assert rel_err(o[:, :, :, sl], ref[:, :, :, sl]) < 1.6e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm(dtype):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4 * 124, 256, generator=g) * 3 + 0.5
    gamma, beta = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g)
    pos = torch.randn(124, 256, generator=g)
    ref = F.layer_norm(x, (256,), gamma, beta, 1e-5)
    y, ypos, y32 = gpu_ops.layernorm(x, gamma, beta, pos, dtype)
    torch.cuda.synchronize()
    tol = 2e-6 if dtype == torch.float32 else 4e-3
    assert rel_err(y32, ref) < 2e-6
    assert rel_err(y, ref) < tol
    assert rel_err(ypos, ref + pos.repeat(4, 1)) < tol


def _mha_core_ref(q, k, v, nheads, kpm=None, amask=None):
    B, Lq, E = q.shape
    Lk = k.shape[1]
    hd = E // nheads
    qh = q.view(B, Lq, nheads, hd).transpose(1, 2) * math.sqrt(1.0 / hd)
    kh = k.view(B, Lk, nheads, hd).transpose(1, 2)
    vh = v.view(B, Lk, nheads, hd).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2)
    if amask is not None:
        s = s + amask
    if kpm is not None:
        s = s.masked_fill(kpm.view(B, 1, 1, Lk), float("-inf"))
    return (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, Lq, E)


@pytest.mark.parametrize("Lq,Lk", [(124, 124), (128, 128), (21, 124), (11, 128), (21, 21), (200, 200)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attention(Lq, Lk, dtype):
    g = torch.Generator().manual_seed(Lq * 1000 + Lk)
    B = 3
    q, k, v = (torch.randn(B, L, 256, generator=g) for L in (Lq, Lk, Lk))
    kpm = torch.zeros(B, Lk, dtype=torch.bool)
    kpm[1, Lk - 17:] = True
    kpm[2, Lk // 2:] = True
    amask = None
    if Lq == Lk == 21:
        amask = sedt_oracle.spsedt_attention_mask(20, 10)
        amask = F.pad(amask, (0, 1, 0, 1), value=0.0)        # 21x21 with an unmasked extra row/col
        kpm = None
    if dtype == torch.bfloat16:
        q, k, v = q.bfloat16().float(), k.bfloat16().float(), v.bfloat16().float()
    ref = _mha_core_ref(q, k, v, 8, kpm, amask)
    out = gpu_ops.attention(q.cuda().to(dtype), k.cuda().to(dtype), v.cuda().to(dtype), 8, kpm, amask)
    torch.cuda.synchronize()
    assert rel_err(out, ref) < (3e-6 if dtype == torch.float32 else 4e-3)


def test_pos_table_unpadded_and_padded():
    for T, H in ((500, 32), (496, 31)):
        W = 4
        ref = sedt_oracle.position_sine(torch.zeros(1, H, W, dtype=torch.bool), 256)     # [1,256,H,W]
        pos = gpu_ops.pos_table(None, 1, T, 64, H, W)
        torch.cuda.synchronize()
        got = pos.view(1, H, W, 256).permute(0, 3, 1, 2).cpu()
        assert (got - ref).abs().max() < 2e-6
    # padded: clips of 500 / 333 / 420 frames in a 500-frame batch
    T, H, W = 500, 32, 4
    mask = torch.ones(3, T, 64, dtype=torch.bool)
    for i, t in enumerate((500, 333, 420)):
        mask[i, :t] = False
    mref = sedt_oracle.resize_mask(mask, (H, W))
    ref = sedt_oracle.position_sine(mref, 256)
    pos, ds = gpu_ops.pos_table(mask, 3, T, 64, H, W)
    torch.cuda.synchronize()
    assert torch.equal(ds.view(3, H, W).cpu().bool(), mref)
    assert (pos.view(3, H, W, 256).permute(0, 3, 1, 2).cpu() - ref).abs().max() < 2e-6
