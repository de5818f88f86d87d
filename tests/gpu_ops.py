"""Thin ctypes callers for the single-operator C-ABI entry points (test helper)."""
import ctypes as C

import torch

from sound_event_detection_transformer_b200 import _lib

DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _out(n):
    return (n + 2 * 0 - 1) // 1 + 1


def conv_out(n, k, stride, pad, dil):
    return (n + 2 * pad - dil * (k - 1) - 1) // stride + 1


def repack(w_oihw: torch.Tensor, dtype) -> torch.Tensor:
    lib = _lib.load()
    co, ci, r, s = w_oihw.shape
    w = w_oihw.cuda().float().contiguous()
    out = torch.empty(co, r, s, ci, dtype=dtype, device="cuda")
    _lib.check(lib.sedt_op_repack_conv(w.data_ptr(), out.data_ptr(), DT[dtype], co, ci, r, s, _lib.current_stream()))
    return out


def conv(x_nhwc, w_krsc, scale=None, bias=None, residual=None, stride=1, dil=1, pad=0, relu=False, out_dtype=None,
         engine=0):
    """x_nhwc [B,H,W,Cin] cuda; w_krsc [Cout,R,S,Cin] same dtype; returns [B,Ho,Wo,Cout]."""
    lib = _lib.load()
    B, H, W, Cin = x_nhwc.shape
    Cout, R, S, _ = w_krsc.shape
    Ho, Wo = conv_out(H, R, stride, pad, dil), conv_out(W, S, stride, pad, dil)
    out_dtype = out_dtype or x_nhwc.dtype
    out = torch.empty(B, Ho, Wo, Cout, dtype=out_dtype, device="cuda")
    x_nhwc = x_nhwc.contiguous()
    w_krsc = w_krsc.contiguous()
    d = _lib.SedtConvDesc(in_=x_nhwc.data_ptr(), w=w_krsc.data_ptr(), scale=_lib.ptr(scale) or None,
                          bias=_lib.ptr(bias) or None, residual=_lib.ptr(residual) or None, out=out.data_ptr(),
                          in_dtype=DT[x_nhwc.dtype], out_dtype=DT[out_dtype], B=B, H=H, W=W, Cin=Cin, lda=Cin, Ho=Ho,
                          Wo=Wo, Cout=Cout, ldc=Cout, ld_res=Cout, R=R, S=S, stride=stride, dil=dil, pad=pad,
                          relu=int(relu))
    _lib.check(lib.sedt_op_conv(C.byref(d), engine, _lib.current_stream()))
    return out


def conv_wgrad(x_nhwc, dy_nhwc, k, stride=1, dil=1, pad=0):
    """x [B,H,W,Cin] bf16, dy [B,Ho,Wo,Cout] bf16 -> dw [Cout,k,k,Cin] fp32."""
    lib = _lib.load()
    B, H, W, Cin = x_nhwc.shape
    Cout = dy_nhwc.shape[-1]
    dw = torch.zeros(Cout, k, k, Cin, dtype=torch.float32, device="cuda")
    x_nhwc, dy_nhwc = x_nhwc.contiguous(), dy_nhwc.contiguous()
    _lib.check(lib.sedt_op_conv_wgrad(x_nhwc.data_ptr(), dy_nhwc.data_ptr(), dw.data_ptr(), B, H, W, Cin, Cout, k, stride, dil,
                                      pad, _lib.current_stream()))
    return dw


def tc_supported(x_nhwc, w_krsc, stride=1, dil=1, pad=0, out_dtype=None) -> bool:
    lib = _lib.load()
    B, H, W, Cin = x_nhwc.shape
    Cout, R, S, _ = w_krsc.shape
    Ho, Wo = conv_out(H, R, stride, pad, dil), conv_out(W, S, stride, pad, dil)
    out_dtype = out_dtype or x_nhwc.dtype
    d = _lib.SedtConvDesc(in_=x_nhwc.data_ptr(), w=w_krsc.data_ptr(), scale=None, bias=None, residual=None,
                          out=x_nhwc.data_ptr(), in_dtype=DT[x_nhwc.dtype], out_dtype=DT[out_dtype], B=B, H=H, W=W,
                          Cin=Cin, lda=Cin, Ho=Ho, Wo=Wo, Cout=Cout, ldc=Cout, ld_res=Cout, R=R, S=S, stride=stride,
                          dil=dil, pad=pad, relu=0)
    return bool(lib.sedt_op_conv_tc_supported(C.byref(d)))


def stem(x, sd, body, out_dtype=torch.float32, engine=0):
    lib = _lib.load()
    B, _, T, F = x.shape
    hc = (T - 1) // 2 + 1
    hp = (hc - 1) // 2 + 1
    g = {k: sd[body + k].cuda().float().contiguous() for k in
         ("conv0.weight", "conv0.bias", "conv1.weight", "bn1.weight", "bn1.bias", "bn1.running_mean", "bn1.running_var")}
    scratch = torch.empty(64 * 1024, dtype=torch.uint8, device="cuda")
    sp = (scratch.data_ptr() + 255) & ~255
    out = torch.empty(B, hp, 16, 64, dtype=out_dtype, device="cuda")
    xx = x.cuda().float().contiguous()
    if engine == 1:
        assert out_dtype == torch.bfloat16
        _lib.check(lib.sedt_op_stem_tc(xx.data_ptr(), g["conv0.weight"].data_ptr(), g["conv0.bias"].data_ptr(),
                                       g["conv1.weight"].data_ptr(), g["bn1.weight"].data_ptr(), g["bn1.bias"].data_ptr(),
                                       g["bn1.running_mean"].data_ptr(), g["bn1.running_var"].data_ptr(), sp, out.data_ptr(),
                                       B, T, F, _lib.current_stream()))
        return out
    _lib.check(lib.sedt_op_stem(xx.data_ptr(), g["conv0.weight"].data_ptr(), g["conv0.bias"].data_ptr(),
                                g["conv1.weight"].data_ptr(), g["bn1.weight"].data_ptr(), g["bn1.bias"].data_ptr(),
                                g["bn1.running_mean"].data_ptr(), g["bn1.running_var"].data_ptr(), sp, out.data_ptr(),
                                DT[out_dtype], B, T, F, _lib.current_stream()))
    return out


def layernorm(x, gamma, beta, pos=None, dtype=torch.float32):
    lib = _lib.load()
    rows = x.numel() // 256
    x = x.cuda().float().contiguous()
    y = torch.empty(rows, 256, dtype=dtype, device="cuda")
    ypos = torch.empty(rows, 256, dtype=dtype, device="cuda") if pos is not None else None
    y32 = torch.empty(rows, 256, dtype=torch.float32, device="cuda")
    g, b = gamma.cuda().float().contiguous(), beta.cuda().float().contiguous()
    p = pos.cuda().float().contiguous() if pos is not None else None
    _lib.check(lib.sedt_op_layernorm(x.data_ptr(), g.data_ptr(), b.data_ptr(), _lib.ptr(p) or None,
                                     (p.numel() // 256) if p is not None else 1, y.data_ptr(), _lib.ptr(ypos) or None,
                                     y32.data_ptr(), DT[dtype], rows, _lib.current_stream()))
    return y, ypos, y32


def attention(q, k, v, nheads=8, kpm=None, amask=None):
    """q [B,Lq,256], k/v [B,Lk,256] cuda (fp32 or bf16) -> [B,Lq,256]"""
    lib = _lib.load()
    B, Lq, E = q.shape
    Lk = k.shape[1]
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    o = torch.empty_like(q)
    kp = kpm.cuda().to(torch.uint8).contiguous() if kpm is not None else None
    am = amask.cuda().float().contiguous() if amask is not None else None
    scale = float((1.0 / (E // nheads)) ** 0.5)
    _lib.check(lib.sedt_op_attention(q.data_ptr(), E, k.data_ptr(), E, v.data_ptr(), E, o.data_ptr(), E, DT[q.dtype],
                                     _lib.ptr(kp) or None, _lib.ptr(am) or None, B, nheads, Lq, Lk, scale,
                                     _lib.current_stream()))
    return o


def pos_table(mask, B, T, F, H, W):
    lib = _lib.load()
    if mask is None:
        pos = torch.empty(1, H * W, 256, device="cuda")
        _lib.check(lib.sedt_op_pos_table(None, None, pos.data_ptr(), 1, T, F, H, W, _lib.current_stream()))
        return pos
    m = mask.cuda().to(torch.uint8).contiguous()
    ds = torch.empty(B, H * W, dtype=torch.uint8, device="cuda")
    pos = torch.empty(B, H * W, 256, device="cuda")
    _lib.check(lib.sedt_op_pos_table(m.data_ptr(), ds.data_ptr(), pos.data_ptr(), B, T, F, H, W, _lib.current_stream()))
    return pos, ds


# ---- backward building blocks -----------------------------------------------------------------
def repack_dgrad(w_oihw: torch.Tensor, scale=None, dtype=torch.bfloat16) -> torch.Tensor:
    """OIHW fp32 -> [Cin, R, S, Cout] (taps rotated by 180 degrees, optional per-Cout scale)."""
    lib = _lib.load()
    co, ci, r, s = w_oihw.shape
    w = w_oihw.cuda().float().contiguous()
    out = torch.empty(ci, r, s, co, dtype=dtype, device="cuda")
    _lib.check(lib.sedt_op_repack_dgrad(w.data_ptr(), _lib.ptr(scale) or None, out.data_ptr(), DT[dtype], co, ci, r, s,
                                        _lib.current_stream()))
    return out


def upsample2(dy_nhwc, H, W):
    lib = _lib.load()
    B, Ho, Wo, Cc = dy_nhwc.shape
    u = torch.empty(B, H, W, Cc, dtype=dy_nhwc.dtype, device="cuda")
    _lib.check(lib.sedt_op_upsample2(dy_nhwc.contiguous().data_ptr(), u.data_ptr(), B, H, W, Ho, Wo, Cc, _lib.current_stream()))
    return u


def conv_dgrad(dy_nhwc, w_oihw, in_hw, scale=None, stride=1, dil=1, mask=None, residual=None, out_dtype=torch.bfloat16, engine=1):
    """Data gradient of conv(x, w) (k in {1,3}, pad = dil for k=3) through the forward kernels."""
    H, W = in_hw
    k = w_oihw.shape[-1]
    wd = repack_dgrad(w_oihw, scale)
    g = dy_nhwc
    if stride == 2:
        g = upsample2(dy_nhwc, H, W)
    pad = dil if k == 3 else 0
    assert mask is None or residual is None
    return conv(g, wd, None, None, mask if mask is not None else residual, 1, dil, pad, 2 if mask is not None else 0, out_dtype,
                engine=engine)


def relu_mask(act, g1, g2=None):
    lib = _lib.load()
    out = torch.empty_like(g1)
    _lib.check(lib.sedt_op_relu_mask(act.data_ptr(), g1.data_ptr(), _lib.ptr(g2) or None, out.data_ptr(), g1.numel(),
                                     _lib.current_stream()))
    return out


def colsum(x2d):
    lib = _lib.load()
    M, N = x2d.shape
    out = torch.zeros(N, dtype=torch.float32, device="cuda")
    _lib.check(lib.sedt_op_colsum(x2d.data_ptr(), DT[x2d.dtype], x2d.stride(0), out.data_ptr(), M, N, _lib.current_stream()))
    return out


def layernorm_bwd(x, gamma, g1=None, g2=None, g3=None, dres=None):
    lib = _lib.load()
    rows = x.shape[0]
    dx = torch.empty_like(x)
    dg = torch.zeros(256, dtype=torch.float32, device="cuda")
    db = torch.zeros(256, dtype=torch.float32, device="cuda")
    _lib.check(lib.sedt_op_layernorm_bwd(x.data_ptr(), gamma.data_ptr(), _lib.ptr(g1) or None, _lib.ptr(g2) or None,
                                         _lib.ptr(g3) or None, _lib.ptr(dres) or None, dx.data_ptr(), dg.data_ptr(),
                                         db.data_ptr(), rows, _lib.current_stream()))
    return dx, dg, db


def attention_bwd(q, k, v, do, nheads, kpm=None, amask=None, scale=None, engine=1):
    """q, do [B, Lq, E]; k, v [B, Lk, E] bf16 -> dq, dk, dv."""
    lib = _lib.load()
    B, Lq, E = q.shape
    Lk = k.shape[1]
    scale = scale if scale is not None else (E // nheads) ** -0.5
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    k8 = kpm.to(torch.uint8).contiguous() if kpm is not None else None
    _lib.check(lib.sedt_op_attention_bwd(q.data_ptr(), E, k.data_ptr(), E, v.data_ptr(), E, do.data_ptr(), E, dq.data_ptr(), E,
                                         dk.data_ptr(), E, dv.data_ptr(), E, _lib.ptr(k8) or None, _lib.ptr(amask) or None, B,
                                         nheads, Lq, Lk, scale, engine, _lib.current_stream()))
    return dq, dk, dv
