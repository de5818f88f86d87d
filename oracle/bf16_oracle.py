"""ORACLE (test infrastructure, not product code).

bf16-rounding restatement of the SEDT forward: the arithmetic of
oracle/sedt_oracle.py (which is pinned to the reference's golden outputs, see
its header) with round-to-nearest-even bf16 inserted at exactly the points where
the bf16 tier of the CUDA path stores bf16:

  * conv / linear weights rounded to bf16, the FrozenBN scale folded into the conv
    weights BEFORE rounding (sedt/backbone.py:43-53 folded as in pack.cu);
  * every backbone activation stored as bf16 (stem output, conv1/conv2/conv3 outputs
    after bias + residual + ReLU); the input clip rounded to bf16 by the stem;
  * the transformer residual stream stays fp32; LayerNorm outputs (LN(x), LN(x)+pos),
    Q/K/V, the un-normalised softmax numerators P, the attention output, the FFN hidden
    activation, encoder memory (and memory+pos), cross-attention K/V, the decoder
    states fed to the box MLP and its first hidden layer are bf16;
  * class / last box / tag projections in fp32 on the fp32 decoder states.

All products are therefore exact in fp32 and only the accumulation order differs
from the kernels.  Rounding is a straight-through op for autograd (identity
gradient), so `sedt_forward_bf16(..., grad=True)` gives the fp32 gradient of the
bf16-rounded forward: ReLU masks, attention weights and LayerNorm statistics are
the kernels' own, which removes the sqrt(flipped-ReLU-fraction) term that a plain
fp32 reference shows (tests/test_gpu_train.py docstring) and allows a ~1e-2 bar.

PARITY STATUS: derived, not independently pinned -- with rounding disabled
(`ROUND = False`) it must reproduce sedt_oracle bit for bit up to fp32 summation
order (tests/test_oracle_golden.py::test_bf16_oracle_reduces_to_fp32_oracle).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

from . import sedt_oracle as so

BODY = so.BODY
ROUND = True


class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


def r16(x):
    return _RoundBF16.apply(x) if ROUND else x


def _bn_fold(sd, name):
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    rv, rm = sd[name + ".running_var"], sd[name + ".running_mean"]
    scale = w * (rv + 1e-5).rsqrt()
    return scale, b - rm * scale


def _conv_bn(x, sd, conv, bn, stride=1, padding=0, dilation=1):
    """bf16 conv weights with the BN scale folded in before rounding; fp32 bias (pack.cu: repack_conv / bn_fold)."""
    scale, bias = _bn_fold(sd, bn)
    w = r16(sd[conv + ".weight"] * scale.view(-1, 1, 1, 1))
    return F.conv2d(x, w, None, stride=stride, padding=padding, dilation=dilation) + bias.view(1, -1, 1, 1)


def _stem(sd, x):
    """stem_tc.cu: conv0 folded into conv1 (one input channel, 7x7, K = 49): bf16 operands; the conv0-bias term only counts
    the taps that fall inside the clip (zero padding happens AFTER conv0 in the reference) and is added in fp32."""
    scale, bias = _bn_fold(sd, BODY + "bn1")
    w0 = sd[BODY + "conv0.weight"].view(3)
    b0 = sd[BODY + "conv0.bias"].view(3)
    w1 = sd[BODY + "conv1.weight"]                                   # [64, 3, 7, 7]
    weff = (w1 * w0.view(1, 3, 1, 1)).sum(1, keepdim=True)           # [64, 1, 7, 7]
    wb = (w1 * b0.view(1, 3, 1, 1)).sum(1, keepdim=True)
    y = F.conv2d(r16(x), r16(weff * scale.view(-1, 1, 1, 1)), None, stride=2, padding=3)
    inside = torch.ones_like(x[:1])
    yb = F.conv2d(inside, wb, None, stride=2, padding=3) * scale.view(1, -1, 1, 1)
    y = F.relu(y + yb + bias.view(1, -1, 1, 1))
    return r16(F.max_pool2d(y, kernel_size=3, stride=2, padding=1))


def _bottleneck(x, sd, p, stride, dilation):
    out = r16(F.relu(_conv_bn(x, sd, p + "conv1", p + "bn1")))
    out = r16(F.relu(_conv_bn(out, sd, p + "conv2", p + "bn2", stride=stride, padding=dilation, dilation=dilation)))
    out = _conv_bn(out, sd, p + "conv3", p + "bn3")
    identity = x
    if (p + "downsample.0.weight") in sd:
        identity = r16(_conv_bn(x, sd, p + "downsample.0", p + "downsample.1", stride=stride))
    return r16(F.relu(out + identity))


def backbone_forward_bf16(sd, x, dilation=True):
    x = _stem(sd, x)
    cur_dil = 1
    for li, (nblocks, stride) in enumerate(((3, 1), (4, 2), (6, 2), (3, 2)), start=1):
        prev_dil = cur_dil
        if li == 4 and dilation:
            cur_dil *= stride
            stride = 1
        for bi in range(nblocks):
            x = _bottleneck(x, sd, f"{BODY}layer{li}.{bi}.", stride if bi == 0 else 1, prev_dil if bi == 0 else cur_dil)
    return x


def _dropbf(site, t):
    """sedt_oracle's DROPOUT hooks are sequence-first [L, B, C]; this file is batch-first."""
    if so.DROPOUT is None:
        return t
    return so.DROPOUT(site, t.transpose(0, 1)).transpose(0, 1)


def _lin(x, sd, name, rows=None, f32=False):
    w, b = sd[name + ".weight" if not name.endswith("in_proj") else name + "_weight"], \
        sd[name + ".bias" if not name.endswith("in_proj") else name + "_bias"]
    if rows is not None:
        w, b = w[rows], b[rows]
    return F.linear(x, w if f32 else r16(w), b)


def _attention(q, k, v, nheads, kpm, attn_mask, drop_site):
    """attention_tc.cu: S = Q K^T in fp32 from bf16 Q, K; p = 2^((s - m) * scale * log2 e) with the row maximum over the
    valid keys; the row sum uses the fp32 p, the P V product their bf16 roundings; O = (P V) / sum, stored bf16.
    q [B, L, E], k/v [B, S, E] (batch first)."""
    B, L, E = q.shape
    S = k.shape[1]
    hd = E // nheads
    qh = q.view(B, L, nheads, hd).transpose(1, 2)
    kh = k.view(B, S, nheads, hd).transpose(1, 2)
    vh = v.view(B, S, nheads, hd).transpose(1, 2)
    s = torch.matmul(qh, kh.transpose(-1, -2)) * math.sqrt(1.0 / hd)
    if attn_mask is not None:
        s = s + attn_mask.view(1, 1, L, S)
    if kpm is not None:
        s = s.masked_fill(kpm.view(B, 1, 1, S), float("-inf"))
    m = s.max(-1, keepdim=True).values
    p = torch.exp(s - m)
    l = p.sum(-1, keepdim=True)
    if drop_site is not None and so.DROPOUT is not None:
        # the reference drops the normalised weights; p/l dropped == dropped(p)/l
        p = so.DROPOUT(drop_site, (p / l).reshape(B * nheads, L, S)).view(B, nheads, L, S) * l
    o = torch.matmul(r16(p), vh) / l
    return r16(o.transpose(1, 2).reshape(B, L, E))


def _mha_self(x_ln, x_lnpos, sd, name, nheads, kpm, attn_mask, site):
    d = x_ln.shape[-1]
    qk = r16(_lin(x_lnpos, sd, name + ".in_proj", slice(0, 2 * d)))
    v = r16(_lin(x_ln, sd, name + ".in_proj", slice(2 * d, 3 * d)))
    ao = _attention(qk[..., :d], qk[..., d:], v, nheads, kpm, attn_mask, site)
    return _lin(ao, sd, name + ".out_proj")


def _ffn(x16, sd, p):
    h = _dropbf(p + "hidden", r16(F.relu(_lin(x16, sd, p + "linear1"))))
    return _lin(h, sd, p + "linear2")


def sedt_forward_bf16(sd, args, samples, mask: Optional[torch.Tensor] = None, grad: bool = False, taps: Optional[dict] = None):
    """Same contract as sedt_oracle.sedt_forward (pre-norm only).  grad=True keeps the autograd graph."""
    assert args.pre_norm, "the bf16 tier's rounding points are restated for the pre-norm layers"
    with torch.set_grad_enabled(grad):
        x, mask = so.nested(samples, mask)
        feat = backbone_forward_bf16(sd, x, args.dilation)
        if taps is not None:
            taps["layer4"] = feat
        m = so.resize_mask(mask, feat.shape[-2:])
        pos = so.position_sine(m, args.hidden_dim).flatten(2).transpose(1, 2)        # [B, S, d]
        kpm = m.flatten(1)
        kpm = kpm if bool(kpm.any()) else None
        B, C, H, W = feat.shape
        d, nh = args.hidden_dim, args.nheads
        tok = feat.flatten(2).transpose(1, 2)                                        # [B, S, 2048] (bf16 values)
        xs = F.linear(tok, r16(sd["input_proj.weight"].view(d, C)), sd["input_proj.bias"])
        for li in range(args.enc_layers):
            p = f"transformer.encoder.layers.{li}."
            ln = so.layer_norm(xs, sd, p + "norm1")
            xs = xs + _dropbf(p + "drop1", _mha_self(r16(ln), r16(ln + pos), sd, p + "self_attn", nh, kpm, None, p + "attn"))
            ln2 = r16(so.layer_norm(xs, sd, p + "norm2"))
            xs = xs + _dropbf(p + "drop2", _ffn(ln2, sd, p))
        mem = so.layer_norm(xs, sd, "transformer.encoder.norm")
        if taps is not None:
            taps["memory"] = mem
        mem16, mempos16 = r16(mem), r16(mem + pos)
        qe = sd["query_embed.weight"]
        Q = qe.shape[0]
        t = torch.zeros(B, Q, d)
        inter = []
        for li in range(args.dec_layers):
            p = f"transformer.decoder.layers.{li}."
            ln = so.layer_norm(t, sd, p + "norm1")
            t = t + _dropbf(p + "drop1", _mha_self(r16(ln), r16(ln + qe), sd, p + "self_attn", nh, None, None, p + "self_attn"))
            ln2 = so.layer_norm(t, sd, p + "norm2")
            name = p + "multihead_attn"
            q = r16(_lin(r16(ln2 + qe), sd, name + ".in_proj", slice(0, d)))
            k = r16(_lin(mempos16, sd, name + ".in_proj", slice(d, 2 * d)))
            v = r16(_lin(mem16, sd, name + ".in_proj", slice(2 * d, 3 * d)))
            ao = _attention(q, k, v, nh, kpm, None, p + "cross_attn")
            t = t + _dropbf(p + "drop2", _lin(ao, sd, name + ".out_proj"))
            ln3 = r16(so.layer_norm(t, sd, p + "norm3"))
            t = t + _dropbf(p + "drop3", _ffn(ln3, sd, p))
            inter.append(so.layer_norm(t, sd, "transformer.decoder.norm"))
        hs = torch.stack(inter)                                                       # [D, B, Q, d] fp32
        if taps is not None:
            taps["hs"] = hs
        start = 1 if args.dec_at else 0
        hq = hs[:, :, start:, :]
        h1 = r16(F.relu(_lin(r16(hq), sd, "bbox_embed.layers.0")))
        h2 = F.relu(_lin(h1, sd, "bbox_embed.layers.1"))
        out = {}
        coord = _lin(h2, sd, "bbox_embed.layers.2", f32=True).sigmoid()
        cls = _lin(hq, sd, "class_embed", f32=True)
        if args.dec_at:
            out["at"] = _lin(hs[-1, :, 0, :], sd, "weak_class_embed", f32=True).squeeze().sigmoid()
        out["pred_logits"], out["pred_boxes"] = cls[-1], coord[-1]
        if args.aux_loss:
            out["aux_outputs"] = [{"pred_logits": a, "pred_boxes": b} for a, b in zip(cls[:-1], coord[:-1])]
        return out
