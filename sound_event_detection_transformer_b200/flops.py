"""Algorithmic FLOPs (2 x MACs) of the SEDT forward per clip, from the layer table.
Cross-checked against FlopCounterMode on the reference (SURVEY.md section 8d:
config 1 = 9.458 G, config 2 = 10.333 G per clip)."""
from __future__ import annotations

from .spec import backbone_blocks, conv_out


def forward_flops_per_clip(args, T: int, F: int = 64) -> dict:
    h, w = conv_out(T, 7, 2, 3), conv_out(F, 7, 2, 3)
    stem = 2 * h * w * 64 * (3 * 49) + 2 * T * F * 3            # conv1 as the reference runs it + conv0
    h, w = conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)
    conv = 0
    for blk in backbone_blocks(args.dilation):
        c2 = blk.conv2
        h2, w2 = conv_out(h, 3, c2.stride, c2.pad, c2.dilation), conv_out(w, 3, c2.stride, c2.pad, c2.dilation)
        conv += 2 * h * w * blk.conv1.cout * blk.conv1.cin                         # 1x1 at the input resolution
        conv += 2 * h2 * w2 * c2.cout * c2.cin * 9                                 # 3x3 carries the stride
        conv += 2 * h2 * w2 * blk.conv3.cout * blk.conv3.cin                       # 1x1 at the output resolution
        if blk.downsample is not None:
            conv += 2 * h2 * w2 * blk.downsample.cout * blk.downsample.cin
        h, w = h2, w2
    S = h * w
    d, ff, nh = args.hidden_dim, args.dim_feedforward, args.nheads
    q = args.num_queries + (1 if args.dec_at else 0)
    input_proj = 2 * S * 2048 * d
    attn = lambda lq, lk: 2 * 2 * lq * lk * d                  # QK^T and PV over all heads
    enc_layer = 2 * S * d * d * 4 + attn(S, S) + 2 * 2 * S * d * ff
    dec_layer = (2 * q * d * d * 4 + attn(q, q)) + (2 * q * d * d * 2 + 2 * S * d * d * 2 + attn(q, S)) + 2 * 2 * q * d * ff
    ncls = 1 if args.self_sup else args.num_classes
    heads = args.dec_layers * (q * 2 * d * (ncls + 1) + q * 2 * (d * d * 2 + d * 2)) + 2 * d * ncls
    enc, dec = args.enc_layers * enc_layer, args.dec_layers * dec_layer
    gemm = conv + input_proj + enc + dec - args.enc_layers * attn(S, S) - args.dec_layers * (attn(q, q) + attn(q, S))
    total = stem + conv + input_proj + enc + dec + heads
    return dict(total=total, stem=stem, backbone=stem + conv, conv=conv, input_proj=input_proj, encoder=enc, decoder=dec,
                heads=heads, tensor_core_gemm=gemm, S=S)
