"""The TMA + tcgen05 implicit-GEMM kernel (engine 1) against torch fp32 conv on the same
bf16-rounded operands, for every geometry the backbone and the transformer use.  GPU only."""
import math

import pytest
import torch
import torch.nn.functional as F

import gpu_ops
from test_gpu_ops import _conv_case, rel_err

pytestmark = pytest.mark.gpu

TC_SHAPES = [
    # B, H, W, Cin, Cout, k, stride, dil, residual, relu          (what it stands for)
    (2, 124, 16, 64, 64, 1, 1, 1, False, True),      # layer1 conv1
    (2, 124, 16, 64, 64, 3, 1, 1, False, True),      # layer1 conv2 (3x3, zero padding via TMA OOB)
    (2, 125, 16, 64, 256, 1, 1, 1, True, True),      # layer1 conv3 + residual, T=500 (partial last tile)
    (2, 124, 16, 256, 128, 1, 1, 1, False, True),    # layer2.0 conv1
    (2, 124, 16, 128, 128, 3, 2, 1, False, True),    # layer2.0 conv2, stride 2 (phase views)
    (2, 125, 16, 128, 128, 3, 2, 1, False, True),    # same with odd H
    (2, 124, 16, 256, 512, 1, 2, 1, False, False),   # layer2.0 downsample, stride 2
    (3, 62, 8, 128, 128, 3, 1, 1, False, True),      # layer2 conv2
    (3, 62, 8, 512, 256, 1, 1, 1, False, True),      # layer3.0 conv1
    (3, 62, 8, 256, 256, 3, 2, 1, False, True),      # layer3.0 conv2 stride 2
    (5, 31, 4, 256, 256, 3, 1, 1, False, True),      # layer3 conv2
    (5, 31, 4, 1024, 512, 1, 1, 1, False, True),     # layer4.0 conv1
    (5, 31, 4, 512, 512, 3, 1, 2, False, True),      # layer4 conv2, dilation 2
    (5, 32, 4, 512, 2048, 1, 1, 1, True, True),      # layer4 conv3 + residual
    (6, 8, 4, 256, 256, 3, 1, 1, False, True),       # SP-SEDT patch at layer3 (4 images per tile)
    (1, 1, 1, 256, 512, 1, 1, 1, False, False),      # degenerate: a single row
    (40, 31, 4, 512, 2048, 1, 1, 1, True, True),     # enough tiles for the BLOCK_N=256 variant, with residual
    (20, 124, 16, 64, 256, 1, 1, 1, True, True),     # BLOCK_N=256, K=64 (one k-block per tile), many tiles per CTA
    (20, 124, 16, 64, 64, 3, 1, 1, False, True),     # BLOCK_N=64, many tiles per CTA
]


@pytest.mark.parametrize("shape", TC_SHAPES)
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_conv_tc(shape, out_dtype):
    B, H, W, Cin, Cout, k, stride, dil, res, relu = shape
    x, w, scale, bias, pad, r, ref = _conv_case(*shape, seed=7, dtype=torch.bfloat16)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    wd = gpu_ops.repack(w, torch.bfloat16)
    assert gpu_ops.tc_supported(xd, wd, stride, dil, pad, out_dtype)
    rd = None
    if res:
        rd = r.permute(0, 2, 3, 1).contiguous().cuda().to(out_dtype)
        if out_dtype == torch.float32:          # the reference used the bf16-rounded residual; keep them equal
            rd = r.permute(0, 2, 3, 1).contiguous().cuda()
    out = gpu_ops.conv(xd, wd, scale.cuda(), bias.cuda(), rd, stride, dil, pad, relu, out_dtype, engine=1)
    torch.cuda.synchronize()
    tol = 1e-5 if out_dtype == torch.float32 else 4e-3       # fp32: accumulation order only; bf16: output rounding
    assert rel_err(out.permute(0, 3, 1, 2), ref) < tol
    # and the two engines agree with each other
    out0 = gpu_ops.conv(xd, wd, scale.cuda(), bias.cuda(), rd, stride, dil, pad, relu, out_dtype, engine=0)
    torch.cuda.synchronize()
    assert rel_err(out, out0) < tol


@pytest.mark.parametrize("rows,K,N", [(248, 256, 512), (1984, 256, 2048), (1984, 2048, 256), (42, 256, 256),
                                      (31744, 2048, 256)])
def test_linear_tc(rows, K, N):
    g = torch.Generator().manual_seed(rows + K + N)
    x = (torch.randn(rows, K, generator=g)).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, generator=g)
    res = torch.randn(rows, N, generator=g)
    ref = F.linear(x.float(), w.float(), b) + res
    out = gpu_ops.conv(x.view(rows, 1, 1, K).cuda(), w.view(N, 1, 1, K).cuda(), None, b.cuda(), res.cuda(),
                       out_dtype=torch.float32, engine=1)
    torch.cuda.synchronize()
    assert rel_err(out.view(rows, N), ref) < 1e-5


def test_conv_tc_inplace_residual():
    """out aliases residual (the transformer's x += f(x) pattern)."""
    g = torch.Generator().manual_seed(5)
    rows, K, N = 372, 2048, 256
    x = torch.randn(rows, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16()
    acc = torch.randn(rows, N, generator=g)
    ref = acc + F.linear(x.float(), w.float())
    import ctypes as C
    from sound_event_detection_transformer_b200 import _lib
    lib = _lib.load()
    xd, wd, ad = x.cuda(), w.cuda(), acc.cuda()
    d = _lib.SedtConvDesc(in_=xd.data_ptr(), w=wd.data_ptr(), scale=None, bias=None, residual=ad.data_ptr(), out=ad.data_ptr(),
                          in_dtype=1, out_dtype=0, B=rows, H=1, W=1, Cin=K, lda=K, Ho=1, Wo=1, Cout=N, ldc=N, ld_res=N,
                          R=1, S=1, stride=1, dil=1, pad=0, relu=0)
    _lib.check(lib.sedt_op_conv(C.byref(d), 1, _lib.current_stream()))
    torch.cuda.synchronize()
    assert rel_err(ad, ref) < 1e-5


TC_2SM_SHAPES = [
    (40, 31, 4, 512, 2048, 1, 1, 1, True, True),      # layer4 conv3 + residual
    (40, 31, 4, 512, 512, 3, 1, 2, False, True),      # layer4 conv2, dilation 2, K = 4608
    (20, 124, 16, 64, 256, 1, 1, 1, True, True),      # layer1 conv3: one k block per tile, many tiles per cluster
    (37, 31, 4, 1024, 256, 1, 1, 1, False, True),     # odd number of M tiles: the peer CTA's last tile is out of range
    (24, 62, 8, 256, 256, 3, 2, 1, False, True),      # stride 2 through phase views
    (300, 1, 124, 256, 512, 1, 1, 1, False, False),   # token-major linear (QK projection shape), no ReLU
]


@pytest.mark.parametrize("shape", TC_2SM_SHAPES)
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_conv_tc_2sm(shape, out_dtype):
    """cta_group::2 kernel (engine 2) against torch fp32 conv and against the CUDA-core kernel."""
    B, H, W, Cin, Cout, k, stride, dil, res, relu = shape
    x, w, scale, bias, pad, r, ref = _conv_case(*shape, seed=9, dtype=torch.bfloat16)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    wd = gpu_ops.repack(w, torch.bfloat16)
    rd = r.permute(0, 2, 3, 1).contiguous().cuda().to(out_dtype) if res else None
    out = gpu_ops.conv(xd, wd, scale.cuda(), bias.cuda(), rd, stride, dil, pad, relu, out_dtype, engine=2)
    torch.cuda.synchronize()
    tol = 4e-3 if out_dtype == torch.bfloat16 else 1e-5
    assert rel_err(out.permute(0, 3, 1, 2), ref) < tol
    out0 = gpu_ops.conv(xd, wd, scale.cuda(), bias.cuda(), rd, stride, dil, pad, relu, out_dtype, engine=0)
    torch.cuda.synchronize()
    assert rel_err(out, out0) < tol


def test_linear_tc_2sm_inplace_residual_f32():
    """FFN2 shape: K = 2048, fp32 out aliased with the fp32 residual (x += f(x))."""
    import ctypes as C
    from sound_event_detection_transformer_b200 import _lib
    g = torch.Generator().manual_seed(15)
    rows, K, N = 31 * 124, 2048, 256
    x = torch.randn(rows, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, generator=g)
    acc = torch.randn(rows, N, generator=g)
    ref = acc + F.linear(x.float(), w.float(), b)
    lib = _lib.load()
    xd, wd, ad, bd = x.cuda(), w.cuda(), acc.cuda(), b.cuda()
    d = _lib.SedtConvDesc(in_=xd.data_ptr(), w=wd.data_ptr(), scale=None, bias=bd.data_ptr(), residual=ad.data_ptr(),
                          out=ad.data_ptr(), in_dtype=1, out_dtype=0, B=rows, H=1, W=1, Cin=K, lda=K, Ho=1, Wo=1, Cout=N,
                          ldc=N, ld_res=N, R=1, S=1, stride=1, dil=1, pad=0, relu=0)
    _lib.check(lib.sedt_op_conv(C.byref(d), 2, _lib.current_stream()))
    torch.cuda.synchronize()
    assert rel_err(ad, ref) < 1e-5


TC_WS_SHAPES = [
    (20, 124, 16, 64, 256, 1, 1, 1, True, True),      # layer1 conv3: K = 64, residual
    (9, 62, 8, 128, 512, 1, 1, 1, True, True),        # layer2 conv3: K = 128
    (37, 31, 4, 256, 1024, 1, 1, 1, True, True),      # layer3 conv3: K = 256, 8 N tiles
    (8, 124, 16, 256, 128, 1, 2, 1, False, False),    # stride-2 1x1 (downsample-like), single N tile
    (3, 1, 1, 256, 2048, 1, 1, 1, False, True),       # fewer M tiles than CTAs per N tile
    (600, 1, 1, 256, 2048, 1, 1, 1, False, True),     # FFN linear1 shape (rows as images)
    (6, 124, 16, 64, 64, 3, 1, 1, False, True),       # layer1 conv2: 3x3, whole 64 x 576 filter resident (bf16 out only)
    (5, 124, 16, 256, 64, 1, 1, 1, False, True),      # layer1 conv1 of blocks 1-2: N = 64, K = 256
]


@pytest.mark.parametrize("shape", TC_WS_SHAPES)
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_conv_tc_weight_stationary(shape, out_dtype):
    """weight-stationary kernel (engine 3) against torch fp32 conv and the CUDA-core kernel."""
    B, H, W, Cin, Cout, k, stride, dil, res, relu = shape
    if Cout == 64 and out_dtype == torch.float32:
        pytest.skip("the 64-wide weight-stationary variant only writes bf16")
    x, w, scale, bias, pad, r, ref = _conv_case(*shape, seed=11, dtype=torch.bfloat16)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda().bfloat16()
    wd = gpu_ops.repack(w, torch.bfloat16)
    rd = r.permute(0, 2, 3, 1).contiguous().cuda().to(out_dtype) if res else None
    out = gpu_ops.conv(xd, wd, scale.cuda(), bias.cuda(), rd, stride, dil, pad, relu, out_dtype, engine=3)
    torch.cuda.synchronize()
    tol = 4e-3 if out_dtype == torch.bfloat16 else 1e-5
    assert rel_err(out.permute(0, 3, 1, 2), ref) < tol
    out0 = gpu_ops.conv(xd, wd, scale.cuda(), bias.cuda(), rd, stride, dil, pad, relu, out_dtype, engine=0)
    torch.cuda.synchronize()
    assert rel_err(out, out0) < tol
