"""ORACLE (test infrastructure, not product code).

CPU fp32 restatement of the reference's SEDT / SP-SEDT forward hot path in
plain torch tensor ops, driven directly by a reference-named state_dict.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this file; the product package never does.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so this restatement is pinned against the reference
itself: tests/golden/make_golden.py imports /root/reference, loads the same
synthetic state_dict with strict=True, runs the reference modules and stores
their outputs under tests/golden/; tests/test_oracle_golden.py checks this
file against those fixtures (measured on the generating machine: stem, every
ResNet stage, encoder memory and decoder hs are bit-identical, diff 0.0; the
head outputs differ by a few ulp and are held to 2e-6 x max(1,|ref|max)).

Each function cites the reference lines it restates (paths relative to
/root/reference, torchvision/torch paths relative to site-packages).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]
BODY = "backbone.0.body."


# --------------------------------------------------------------------------
# backbone
# --------------------------------------------------------------------------
def frozen_bn(x: Tensor, sd: SD, name: str) -> Tensor:
    """sedt/backbone.py:43-53 (FrozenBatchNorm2d.forward, eps=1e-5)."""
    w = sd[name + ".weight"].reshape(1, -1, 1, 1)
    b = sd[name + ".bias"].reshape(1, -1, 1, 1)
    rv = sd[name + ".running_var"].reshape(1, -1, 1, 1)
    rm = sd[name + ".running_mean"].reshape(1, -1, 1, 1)
    scale = w * (rv + 1e-5).rsqrt()
    bias = b - rm * scale
    return x * scale + bias


def bottleneck(x: Tensor, sd: SD, p: str, stride: int, dilation: int) -> Tensor:
    """torchvision/models/resnet.py:143-163 (Bottleneck.forward, v1.5)."""
    identity = x
    out = F.conv2d(x, sd[p + "conv1.weight"])
    out = F.relu(frozen_bn(out, sd, p + "bn1"))
    out = F.conv2d(out, sd[p + "conv2.weight"], stride=stride, padding=dilation, dilation=dilation)
    out = F.relu(frozen_bn(out, sd, p + "bn2"))
    out = F.conv2d(out, sd[p + "conv3.weight"])
    out = frozen_bn(out, sd, p + "bn3")
    if (p + "downsample.0.weight") in sd:
        identity = frozen_bn(F.conv2d(x, sd[p + "downsample.0.weight"], stride=stride), sd, p + "downsample.1")
    return F.relu(out + identity)


def backbone_forward(sd: SD, x: Tensor, dilation: bool = True, taps: Optional[dict] = None) -> Tensor:
    """sedt/backbone.py:97-111 + torchvision resnet.py:225-262,266-281:
    conv0 -> conv1 -> bn1 -> relu -> maxpool -> layer1..4.  x: [N,1,T,F]."""
    x = F.conv2d(x, sd[BODY + "conv0.weight"], sd[BODY + "conv0.bias"])
    x = F.conv2d(x, sd[BODY + "conv1.weight"], stride=2, padding=3)
    x = F.relu(frozen_bn(x, sd, BODY + "bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    if taps is not None:
        taps["stem"] = x
    cur_dil = 1
    for li, (nblocks, stride) in enumerate(((3, 1), (4, 2), (6, 2), (3, 2)), start=1):
        prev_dil = cur_dil
        if li == 4 and dilation:            # resnet.py:235-238
            cur_dil *= stride
            stride = 1
        for bi in range(nblocks):
            x = bottleneck(x, sd, f"{BODY}layer{li}.{bi}.", stride if bi == 0 else 1,
                           prev_dil if bi == 0 else cur_dil)
        if taps is not None:
            taps[f"layer{li}"] = x
    return x


def resize_mask(mask: Tensor, hw: Tuple[int, int]) -> Tensor:
    """sedt/backbone.py:81 (nearest interpolate of the padding mask)."""
    return F.interpolate(mask[None].float(), size=hw).to(torch.bool)[0]


def position_sine(mask: Tensor, num_pos_feats: int = 256, temperature: float = 10000.0) -> Tensor:
    """sedt/position_encoding.py:28-47 (time axis only, normalize=True,
    scale=2*pi).  mask: [B,H,W] bool, True = padding.  Returns [B,C,H,W]."""
    not_mask = ~mask
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    y_embed = y_embed / (y_embed[:, -1:, :] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
    pos_y = y_embed[:, :, :, None] / dim_t
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return pos_y.permute(0, 3, 1, 2)


# --------------------------------------------------------------------------
# transformer
# --------------------------------------------------------------------------
# Train-mode dropout points of the pre-norm layers (transformer.py:192-204, :263-284; functional.py:6650).  None = eval.
# When set: callable(site: str, t: Tensor) -> Tensor applied at "<layer prefix>{attn|drop1|hidden|drop2}" (encoder) and
# "<layer prefix>{self_attn|drop1|cross_attn|drop2|hidden|drop3}" (decoder); tests inject the kernels' own masks.
DROPOUT = None


def _drop(site: str, t: Tensor) -> Tensor:
    return t if DROPOUT is None else DROPOUT(site, t)


def layer_norm(x: Tensor, sd: SD, name: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def mha(query: Tensor, key: Tensor, value: Tensor, sd: SD, name: str, nheads: int,
        attn_mask: Optional[Tensor] = None, key_padding_mask: Optional[Tensor] = None, drop_site: Optional[str] = None) -> Tensor:
    """torch/nn/functional.py:5833-5867 (packed in-proj split into three
    linears because q is k but k is not v) and :6630-6659 (need_weights=True
    eager path: q*sqrt(1/hd), baddbmm/bmm, softmax, bmm, out-proj).
    Shapes are sequence-first: query [L,B,E], key/value [S,B,E]."""
    L, B, E = query.shape
    S = key.shape[0]
    hd = E // nheads
    w, b = sd[name + ".in_proj_weight"], sd[name + ".in_proj_bias"]
    wq, wk, wv = w.chunk(3)
    bq, bk, bv = b.chunk(3)
    q = F.linear(query, wq, bq)
    k = F.linear(key, wk, bk)
    v = F.linear(value, wv, bv)
    q = q.view(L, B * nheads, hd).transpose(0, 1)
    k = k.view(S, B * nheads, hd).transpose(0, 1)
    v = v.view(S, B * nheads, hd).transpose(0, 1)
    mask = None
    if attn_mask is not None:                 # additive float [L,S]
        mask = attn_mask.unsqueeze(0) if attn_mask.dim() == 2 else attn_mask
    if key_padding_mask is not None:          # [B,S] bool, True = ignore -> additive -inf
        kpm = torch.zeros(B, S, dtype=q.dtype).masked_fill_(key_padding_mask, float("-inf"))
        kpm = kpm.view(B, 1, 1, S).expand(-1, nheads, -1, -1).reshape(B * nheads, 1, S)
        mask = kpm if mask is None else mask + kpm
    q_scaled = q * math.sqrt(1.0 / float(hd))
    if mask is not None:
        attn = torch.baddbmm(mask, q_scaled, k.transpose(-2, -1))
    else:
        attn = torch.bmm(q_scaled, k.transpose(-2, -1))
    attn = F.softmax(attn, dim=-1)
    if drop_site is not None:
        attn = _drop(drop_site, attn)
    out = torch.bmm(attn, v)
    out = out.transpose(0, 1).contiguous().view(L * B, E)
    out = F.linear(out, sd[name + ".out_proj.weight"], sd[name + ".out_proj.bias"])
    return out.view(L, B, E)


def ffn(x: Tensor, sd: SD, p: str) -> Tensor:
    return F.linear(_drop(p + "hidden", F.relu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))),
                    sd[p + "linear2.weight"], sd[p + "linear2.bias"])


def encoder_layer(src: Tensor, sd: SD, p: str, nheads: int, pos: Tensor, kpm: Optional[Tensor],
                  pre_norm: bool = True) -> Tensor:
    """sedt/transformer.py:192-204 (forward_pre) / :177-190 (forward_post), eval mode."""
    if pre_norm:
        src2 = layer_norm(src, sd, p + "norm1")
        q = k = src2 + pos
        src = src + _drop(p + "drop1", mha(q, k, src2, sd, p + "self_attn", nheads, key_padding_mask=kpm, drop_site=p + "attn"))
        src2 = layer_norm(src, sd, p + "norm2")
        return src + _drop(p + "drop2", ffn(src2, sd, p))
    q = k = src + pos
    src = layer_norm(src + mha(q, k, src, sd, p + "self_attn", nheads, key_padding_mask=kpm), sd, p + "norm1")
    return layer_norm(src + ffn(src, sd, p), sd, p + "norm2")


def decoder_layer(tgt: Tensor, memory: Tensor, sd: SD, p: str, nheads: int, pos: Tensor, query_pos: Tensor,
                  kpm: Optional[Tensor], tgt_mask: Optional[Tensor], pre_norm: bool = True) -> Tensor:
    """sedt/transformer.py:263-284 (forward_pre) / :240-261 (forward_post), eval mode."""
    if pre_norm:
        t2 = layer_norm(tgt, sd, p + "norm1")
        q = k = t2 + query_pos
        tgt = tgt + _drop(p + "drop1", mha(q, k, t2, sd, p + "self_attn", nheads, attn_mask=tgt_mask, drop_site=p + "self_attn"))
        t2 = layer_norm(tgt, sd, p + "norm2")
        tgt = tgt + _drop(p + "drop2", mha(t2 + query_pos, memory + pos, memory, sd, p + "multihead_attn", nheads,
                                           key_padding_mask=kpm, drop_site=p + "cross_attn"))
        t2 = layer_norm(tgt, sd, p + "norm3")
        return tgt + _drop(p + "drop3", ffn(t2, sd, p))
    q = k = tgt + query_pos
    tgt = layer_norm(tgt + mha(q, k, tgt, sd, p + "self_attn", nheads, attn_mask=tgt_mask), sd, p + "norm1")
    tgt = layer_norm(tgt + mha(tgt + query_pos, memory + pos, memory, sd, p + "multihead_attn", nheads,
                               key_padding_mask=kpm), sd, p + "norm2")
    return layer_norm(tgt + ffn(tgt, sd, p), sd, p + "norm3")


def transformer_forward(sd: SD, src: Tensor, mask: Tensor, query_embed: Tensor, pos: Tensor, *, nheads: int,
                        enc_layers: int, dec_layers: int, pre_norm: bool = True,
                        tgt_mask: Optional[Tensor] = None, query_is_batched: bool = False,
                        taps: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """sedt/transformer.py:48-86 (both branches), :98-111, :123-152.
    src [B,C,H,W] (already input_proj'ed), mask [B,H,W], pos [B,C,H,W].
    query_embed: [Q,C] (supervised) or [Q,B,C] (self_sup branch).
    Returns hs [D,B,Q,C] and memory [B,S,C]."""
    bs = src.shape[0]
    src = src.flatten(2).permute(2, 0, 1)
    pos = pos.flatten(2).permute(2, 0, 1)
    if not query_is_batched:
        query_embed = query_embed.unsqueeze(1).repeat(1, bs, 1)
    kpm = mask.flatten(1)
    tgt = torch.zeros_like(query_embed)
    out = src
    for li in range(enc_layers):
        out = encoder_layer(out, sd, f"transformer.encoder.layers.{li}.", nheads, pos, kpm, pre_norm)
        if taps is not None:
            taps[f"enc{li}"] = out
    if pre_norm:
        out = layer_norm(out, sd, "transformer.encoder.norm")
    memory = out
    inter: List[Tensor] = []
    out = tgt
    for li in range(dec_layers):
        out = decoder_layer(out, memory, sd, f"transformer.decoder.layers.{li}.", nheads, pos, query_embed,
                            kpm, tgt_mask, pre_norm)
        inter.append(layer_norm(out, sd, "transformer.decoder.norm"))
    hs = torch.stack(inter)
    return hs.transpose(1, 2), memory.permute(1, 0, 2)


# --------------------------------------------------------------------------
# model heads
# --------------------------------------------------------------------------
def mlp(x: Tensor, sd: SD, name: str, num_layers: int) -> Tensor:
    """sedt/sedt.py:398-409."""
    for i in range(num_layers):
        x = F.linear(x, sd[f"{name}.layers.{i}.weight"], sd[f"{name}.layers.{i}.bias"])
        if i < num_layers - 1:
            x = F.relu(x)
    return x


def nested(samples, mask: Optional[Tensor]):
    """utilities/utils.py:470-492 for a list of [1,T,F] clips or a [B,1,T,F] tensor."""
    if mask is not None:
        return samples, mask
    if isinstance(samples, torch.Tensor):
        samples = list(samples)
    T = max(s.shape[1] for s in samples)
    Fq = max(s.shape[2] for s in samples)
    x = torch.zeros(len(samples), samples[0].shape[0], T, Fq, dtype=samples[0].dtype)
    m = torch.ones(len(samples), T, Fq, dtype=torch.bool)
    for i, s in enumerate(samples):
        x[i, :, : s.shape[1], : s.shape[2]].copy_(s)
        m[i, : s.shape[1], : s.shape[2]] = False
    return x, m


@torch.no_grad()
def sedt_forward(sd: SD, args, samples, mask: Optional[Tensor] = None, taps: Optional[dict] = None) -> dict:
    """sedt/sedt.py:64-123 (SEDT.forward, dec_at and plain branches; pooling=None)."""
    x, mask = nested(samples, mask)
    feat = backbone_forward(sd, x, args.dilation, taps)
    m = resize_mask(mask, feat.shape[-2:])
    pos = position_sine(m, args.hidden_dim)
    src = F.conv2d(feat, sd["input_proj.weight"], sd["input_proj.bias"])
    hs, memory = transformer_forward(sd, src, m, sd["query_embed.weight"], pos, nheads=args.nheads,
                                     enc_layers=args.enc_layers, dec_layers=args.dec_layers,
                                     pre_norm=args.pre_norm, taps=taps)
    if taps is not None:
        taps["memory"] = memory
        taps["hs"] = hs
    out = {}
    if args.dec_at:
        outputs_class = F.linear(hs[:, :, 1:, :], sd["class_embed.weight"], sd["class_embed.bias"])
        outputs_coord = mlp(hs[:, :, 1:, :], sd, "bbox_embed", 3).sigmoid()
        out["at"] = F.linear(hs[-1, :, 0, :], sd["weak_class_embed.weight"], sd["weak_class_embed.bias"]).squeeze().sigmoid()
    else:
        outputs_class = F.linear(hs, sd["class_embed.weight"], sd["class_embed.bias"])
        outputs_coord = mlp(hs, sd, "bbox_embed", 3).sigmoid()
    out["pred_logits"] = outputs_class[-1]
    out["pred_boxes"] = outputs_coord[-1]
    if args.aux_loss:
        out["aux_outputs"] = [{"pred_logits": a, "pred_boxes": b}
                              for a, b in zip(outputs_class[:-1], outputs_coord[:-1])]
    return out


def spsedt_attention_mask(num_queries: int, num_patches: int) -> Tensor:
    """sedt/spsedt.py:27-32."""
    qpp = num_queries // num_patches
    m = torch.ones(num_queries, num_queries) * float("-inf")
    for i in range(num_patches):
        m[i * qpp:(i + 1) * qpp, i * qpp:(i + 1) * qpp] = 0
    return m


@torch.no_grad()
def spsedt_forward(sd: SD, args, x: Tensor, mask: Tensor, patches: Tensor, taps: Optional[dict] = None,
                   query_keep: Optional[Tensor] = None) -> dict:
    """sedt/spsedt.py:34-91, query_shuffle=False: the eval branch (:70-75) or, with query_keep [B, Q] (1 = keep; the
    reference draws torch.rand(Q, bs, 1) > mask_ratio at :65), the training branch (:63-69):
    decoder_input = query_embed; decoder_input += patches_feature * mask + decoder_input, i.e. 2 * query_embed + mask * patch."""
    feat = backbone_forward(sd, x, args.dilation, taps)
    m = resize_mask(mask, feat.shape[-2:])
    pos = position_sine(m, args.hidden_dim)
    bs, P = patches.shape[0], patches.shape[1]
    pf = backbone_forward(sd, patches.flatten(0, 1), args.dilation)
    gt = F.adaptive_avg_pool2d(pf, (1, 1)).flatten(1)
    qpp = args.num_queries // args.num_patches
    pq = F.linear(gt, sd["patch2query.weight"], sd["patch2query.bias"]).view(bs, P, 1, -1) \
        .repeat(1, 1, qpp, 1).flatten(1, 2).permute(1, 0, 2).contiguous()
    start = 1 if args.dec_at else 0
    nq = P * args.num_queries // args.num_patches
    if query_keep is not None:
        assert nq == args.num_queries - start, "the training branch uses num_patches patches"
        qe = sd["query_embed.weight"][start:, :].unsqueeze(1).repeat(1, bs, 1)
        dec_in = qe + (pq * query_keep.t().to(pq.dtype).unsqueeze(-1) + qe)
    else:
        dec_in = pq + sd["query_embed.weight"][start:nq, :].unsqueeze(1).repeat(1, bs, 1)
    tgt_mask = spsedt_attention_mask(args.num_queries, args.num_patches)[:nq, :nq]
    src = F.conv2d(feat, sd["input_proj.weight"], sd["input_proj.bias"])
    hs, memory = transformer_forward(sd, src, m, dec_in, pos, nheads=args.nheads, enc_layers=args.enc_layers,
                                     dec_layers=args.dec_layers, pre_norm=args.pre_norm, tgt_mask=tgt_mask,
                                     query_is_batched=True, taps=taps)
    outputs_class = F.linear(hs, sd["class_embed.weight"], sd["class_embed.bias"])
    outputs_coord = mlp(hs, sd, "bbox_embed", 3).sigmoid()
    out = {"pred_logits": outputs_class[-1], "pred_boxes": outputs_coord[-1]}
    if args.feature_recon:
        outputs_feature = mlp(hs, sd, "feature_align", 2)
        out["pred_feature"] = outputs_feature[-1]
        out["gt_feature"] = gt
        if args.aux_loss:
            out["aux_outputs"] = [{"pred_logits": a, "pred_boxes": b, "pred_feature": c, "gt_feature": gt}
                                  for a, b, c in zip(outputs_class[:-1], outputs_coord[:-1], outputs_feature[:-1])]
    elif args.aux_loss:
        out["aux_outputs"] = [{"pred_logits": a, "pred_boxes": b}
                              for a, b in zip(outputs_class[:-1], outputs_coord[:-1])]
    return out


@torch.no_grad()
def post_process(outputs: dict, target_sizes: Tensor, audio_tags: Optional[Tensor] = None, at_m: int = 2,
                 threshold: float = 0.5) -> List[dict]:
    """sedt/sedt.py:359-396 (PostProcess.forward, is_semi=False)."""
    out_logits, out_bbox = outputs["pred_logits"], outputs["pred_boxes"]
    bs, num_q, _ = out_logits.shape
    prob = F.softmax(out_logits, -1)
    if audio_tags is not None:
        _, idx = prob[..., :-1].max(1)
        if at_m == 1:
            prob[..., :-1] = prob[..., :-1] * audio_tags.unsqueeze(1).repeat(1, num_q, 1)
        if at_m == 2:
            tags = audio_tags.unsqueeze(1).repeat(1, num_q, 1)
            for i, j in enumerate(idx):
                ar = torch.arange(len(j))
                ind = prob[i, j, ar] < threshold
                prob[i, j[ind], ar[ind]] = threshold
            prob[..., :-1] = prob[..., :-1] * tags
        if at_m == 3:
            for i, (j, at) in enumerate(zip(idx, audio_tags)):
                ar = torch.arange(len(j))
                ind = (prob[i, j, ar] < threshold) & at.bool()
                prob[i, j[ind], ar[ind]] = threshold
    scores, labels = prob[..., :-1].max(-1)
    c, l = out_bbox.unbind(-1)
    boxes = torch.stack([c - l / 2, c + l / 2], dim=-1)        # utilities/box_ops.py:16-19
    boxes = boxes * target_sizes.unsqueeze(-1)[:, None, :]
    return [{"scores": s, "labels": lb, "boxes": b} for s, lb, b in zip(scores, labels, boxes)]
