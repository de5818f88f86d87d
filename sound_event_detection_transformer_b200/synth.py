"""Seeded synthetic weights and inputs for the SEDT hot path.

There is no network for checkpoints, so every parity test, the smoke test and
bench.py use random-init weights of the reference architecture.  The state
dict produced here has exactly the reference's names and shapes
(374 entries for E=3, 410 for E=6) and loads with
`strict=True` into `sedt.build_model(args)[0]` of the reference
(tests/golden/make_golden.py asserts that).

Distributions follow what the reference ends up with at construction
 with two deliberate changes that make parity
checks meaningful:
  * FrozenBatchNorm2d buffers are randomised (the reference constructs the
    identity transform and only ever gets real statistics from a checkpoint);
    the last BN of each bottleneck is scaled down so activations stay O(1)
    through the un-normalised residual trunk instead of growing ~50x;
  * conv weights use He-normal so every stage keeps unit-scale activations.
Both only change numbers, never shapes or names.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from .spec import BODY, backbone_blocks


def _randn(g, *shape, std=1.0):
    return torch.randn(*shape, generator=g, dtype=torch.float32) * std


def _uniform(g, *shape, lo=0.0, hi=1.0):
    return torch.rand(*shape, generator=g, dtype=torch.float32) * (hi - lo) + lo


def _xavier(g, out_f, in_f):
    a = math.sqrt(6.0 / (in_f + out_f))
    return _uniform(g, out_f, in_f, lo=-a, hi=a)


def _linear(g, sd, name, out_f, in_f, w_scale=1.0):
    bound = 1.0 / math.sqrt(in_f)
    sd[name + ".weight"] = _uniform(g, out_f, in_f, lo=-bound, hi=bound) * w_scale
    sd[name + ".bias"] = _uniform(g, out_f, lo=-bound, hi=bound)


def _bn(g, sd, name, n, gain=1.0):
    sd[name + ".weight"] = _uniform(g, n, lo=0.5, hi=1.5) * gain
    sd[name + ".bias"] = _randn(g, n, std=0.1)
    sd[name + ".running_mean"] = _randn(g, n, std=0.1)
    sd[name + ".running_var"] = _uniform(g, n, lo=0.5, hi=1.5)


def _conv(g, sd, name, cout, cin, k):
    fan_in = cin * k * k
    sd[name + ".weight"] = _randn(g, cout, cin, k, k, std=math.sqrt(2.0 / fan_in))


def _mha(g, sd, name, d):
    sd[name + ".in_proj_weight"] = _xavier(g, 3 * d, d)
    sd[name + ".in_proj_bias"] = _randn(g, 3 * d, std=0.02)
    sd[name + ".out_proj.weight"] = _xavier(g, d, d)
    sd[name + ".out_proj.bias"] = _randn(g, d, std=0.02)


def _ln(g, sd, name, d):
    sd[name + ".weight"] = _uniform(g, d, lo=0.8, hi=1.2)
    sd[name + ".bias"] = _randn(g, d, std=0.05)


def synth_state_dict(args, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Reference-named fp32 CPU state dict for `build_model(args)`'s model."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    d, ff = args.hidden_dim, args.dim_feedforward

    for li in range(args.enc_layers):
        p = f"transformer.encoder.layers.{li}."
        _mha(g, sd, p + "self_attn", d)
        sd[p + "linear1.weight"] = _xavier(g, ff, d); sd[p + "linear1.bias"] = _randn(g, ff, std=0.02)
        sd[p + "linear2.weight"] = _xavier(g, d, ff); sd[p + "linear2.bias"] = _randn(g, d, std=0.02)
        _ln(g, sd, p + "norm1", d); _ln(g, sd, p + "norm2", d)
    if args.pre_norm:
        _ln(g, sd, "transformer.encoder.norm", d)
    for li in range(args.dec_layers):
        p = f"transformer.decoder.layers.{li}."
        _mha(g, sd, p + "self_attn", d)
        _mha(g, sd, p + "multihead_attn", d)
        sd[p + "linear1.weight"] = _xavier(g, ff, d); sd[p + "linear1.bias"] = _randn(g, ff, std=0.02)
        sd[p + "linear2.weight"] = _xavier(g, d, ff); sd[p + "linear2.bias"] = _randn(g, d, std=0.02)
        _ln(g, sd, p + "norm1", d); _ln(g, sd, p + "norm2", d); _ln(g, sd, p + "norm3", d)
    _ln(g, sd, "transformer.decoder.norm", d)

    num_classes = 1 if args.self_sup else args.num_classes
    # heads are scaled up so outputs depend visibly on the input
    _linear(g, sd, "class_embed", num_classes + 1, d, w_scale=4.0)
    _linear(g, sd, "bbox_embed.layers.0", d, d, w_scale=2.0)
    _linear(g, sd, "bbox_embed.layers.1", d, d, w_scale=2.0)
    _linear(g, sd, "bbox_embed.layers.2", 2, d, w_scale=4.0)
    sd["input_proj.weight"] = _randn(g, d, 2048, 1, 1, std=math.sqrt(1.0 / 2048))
    sd["input_proj.bias"] = _randn(g, d, std=0.02)

    # backbone (sedt/backbone.py:97-111)
    sd[BODY + "conv0.weight"] = _randn(g, 3, 1, 1, 1, std=1.0)
    sd[BODY + "conv0.bias"] = _randn(g, 3, std=0.3)
    _conv(g, sd, BODY + "conv1", 64, 3, 7)
    _bn(g, sd, BODY + "bn1", 64)
    for blk in backbone_blocks(args.dilation):
        for cs, gain in ((blk.conv1, 1.0), (blk.conv2, 1.0), (blk.conv3, 0.35)):
            _conv(g, sd, cs.name, cs.cout, cs.cin, cs.k)
            _bn(g, sd, cs.bn, cs.cout, gain)
        if blk.downsample is not None:
            cs = blk.downsample
            _conv(g, sd, cs.name, cs.cout, cs.cin, cs.k)
            _bn(g, sd, cs.bn, cs.cout, 0.7)

    nq = args.num_queries + (1 if args.dec_at else 0)
    sd["query_embed.weight"] = _randn(g, nq, d, std=1.0)
    if args.dec_at:
        _linear(g, sd, "weak_class_embed", num_classes, d, w_scale=4.0)
    if args.self_sup:
        _linear(g, sd, "patch2query", d, 2048)
        if args.feature_recon:
            _linear(g, sd, "feature_align.layers.0", d, d)
            _linear(g, sd, "feature_align.layers.1", 2048, d)
    return sd


def synth_clips(B: int, T: int, F: int = 64, seed: int = 0) -> torch.Tensor:
    """Standardised log-mel stand-in: N(0,1) `[B,1,T,F]` fp32."""
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.randn(B, 1, T, F, generator=g, dtype=torch.float32)


def synth_patches(B: int, P: int, PT: int = 128, F: int = 64, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(2000 + seed)
    return torch.randn(B, P, 1, PT, F, generator=g, dtype=torch.float32)


def synth_matcher_case(B: int, Q: int, C: int, kmin: int = 0, kmax: int = 10, seed: int = 3):
    """Config-3 style matcher inputs: continuous => tie-free a.s."""
    g = torch.Generator().manual_seed(3000 + seed)
    logits = torch.randn(B, Q, C + 1, generator=g)
    scale = torch.tensor([1.0, 0.5])
    boxes = torch.rand(B, Q, 2, generator=g) * scale
    sizes = torch.randint(kmin, kmax + 1, (B,), generator=g)
    targets = []
    for k in sizes.tolist():
        targets.append({"labels": torch.randint(0, C, (k,), generator=g),
                        "boxes": torch.rand(k, 2, generator=g) * scale})
    return {"pred_logits": logits, "pred_boxes": boxes}, targets


def synth_criterion_case(B: int, Q: int, C: int, D: int, kmin: int = 0, kmax: int = 10, seed: int = 7):
    """Model-shaped outputs of all D decoder layers (top + D-1 aux) and config-3 style targets for SetCriterion."""
    g = torch.Generator().manual_seed(7000 + seed)
    scale = torch.tensor([1.0, 0.5])

    def layer():
        return {"pred_logits": torch.randn(B, Q, C + 1, generator=g),
                "pred_boxes": (torch.rand(B, Q, 2, generator=g) * 0.9 + 0.05) * scale}
    outputs = layer()
    outputs["at"] = torch.sigmoid(torch.randn(B, C, generator=g))
    outputs["aux_outputs"] = [layer() for _ in range(D - 1)]
    sizes = torch.randint(kmin, kmax + 1, (B,), generator=g)
    targets = []
    for k in sizes.tolist():
        targets.append({"labels": torch.randint(0, C, (k,), generator=g),
                        "boxes": (torch.rand(k, 2, generator=g) * 0.9 + 0.05) * scale,
                        "orig_size": torch.tensor(10.0)})
    return outputs, targets


def synth_teacher_case(B: int, Q: int, C: int, seed: int):
    """Teacher-shaped outputs for the pseudo-label path (same generator as tests/golden/make_golden.py)."""
    g = torch.Generator().manual_seed(9000 + seed)
    logits = torch.randn(B, Q, C + 1, generator=g)
    hot = torch.randint(0, 4, (B, Q), generator=g)
    logits.scatter_add_(2, hot.unsqueeze(-1), ((torch.rand(B, Q, generator=g) > 0.4).float() * 6.0).unsqueeze(-1))
    boxes = torch.stack([torch.rand(B, Q, generator=g), torch.rand(B, Q, generator=g) * 0.3], -1)
    boxes[:, ::5, 1] *= 0.05
    at = torch.rand(B, C, generator=g)
    return logits, boxes, at


def synth_db_clips(lengths, F: int, seed: int):
    """dB-domain features of ragged lengths (same generator as tests/golden/make_golden.py: run_prepare)."""
    g = torch.Generator().manual_seed(9500 + seed)
    return [(torch.randn(t, F, generator=g) * 12.0 - 40.0).numpy().astype("float32") for t in lengths]


def synth_decode_cases(n: int, Q: int, seed: int):
    """Overlap-heavy PostProcess-shaped results (same generator as tests/golden/make_golden.py: run_decode_chains)."""
    import numpy as np
    rng = np.random.default_rng(9700 + seed)
    cases = []
    for _ in range(n):
        onset = rng.uniform(0.0, 8.0, Q).astype(np.float32)
        dur = rng.uniform(0.05, 3.0, Q).astype(np.float32)
        cases.append({"scores": rng.uniform(0.3, 1.0, Q).astype(np.float32), "labels": rng.integers(0, 3, Q).astype(np.int64),
                      "boxes": np.stack([onset, onset + dur], -1).astype(np.float32)})
    return cases


def synth_patch_boxes(B: int, P: int, seed: int, fixed_len=None):
    """(center, width) boxes that stay inside the clip, as SedData.get_random_patch draws them (data_utils/DataLoad.py:57-77):
    widths in [0.0005, 0.8) (or fixed_len), plus the degenerate / full-width cases at the front of clip 0."""
    g = torch.Generator().manual_seed(9800 + seed)
    l = torch.rand(B, P, generator=g) * 0.8 + 0.0005 if fixed_len is None else torch.full((B, P), float(fixed_len))
    c = l / 2 + torch.rand(B, P, generator=g) * (1 - l)
    boxes = torch.stack([c, l], -1)
    if fixed_len is None and P >= 4:
        boxes[0, :4] = torch.tensor([[0.3, 0.0005], [0.5, 1.0], [0.5, 0.26], [0.129, 0.258]])
    return boxes


def synth_mixup_case(n_strong: int, n_weak: int, n_unl: int, T: int, F: int, seed: int, C: int = 10):
    """A batch laid out [strong | weak | unlabelled] like the semi-supervised loader (train_ss_sedt.py): clips + label dicts.
    Strong clips carry 0..12 events (so that > max_events and same-class overlaps occur), weak / unlabelled ones only tags."""
    g = torch.Generator().manual_seed(9900 + seed)
    bs = n_strong + n_weak + n_unl
    x = torch.randn(bs, 1, T, F, generator=g)
    y = []
    for i in range(bs):
        if i < n_strong:
            k = int(torch.randint(0, 13, (1,), generator=g))
            l = torch.rand(k, generator=g) * 0.15 + 0.01
            c = l / 2 + torch.rand(k, generator=g) * (1 - l)
            y.append({"labels": torch.randint(0, C, (k,), generator=g), "boxes": torch.stack([c, l], -1).reshape(k, 2),
                      "orig_size": torch.tensor(10.0)})
        else:
            k = int(torch.randint(0, 3, (1,), generator=g)) if i < n_strong + n_weak else 0
            y.append({"labels": torch.randint(0, C, (k,), generator=g), "boxes": torch.zeros(0, 2), "orig_size": torch.tensor(10.0)})
    return x, y


# ---- SP-SEDT pretraining step: the parameters whose gradients the golden fixture samples, and the functional it differentiates
SP_TRAIN_PARAMS = ("patch2query.weight", "patch2query.bias", "query_embed.weight", "feature_align.layers.1.weight",
                   "feature_align.layers.0.bias", "class_embed.weight", "bbox_embed.layers.0.weight", "input_proj.weight",
                   "transformer.decoder.layers.0.self_attn.in_proj_weight", "transformer.encoder.layers.1.linear1.weight")


def sp_train_functional(out, B, Q, seed):
    """A fixed random linear functional of every output of the pretraining forward (all decoder layers)."""
    g = torch.Generator().manual_seed(8800 + seed)
    tot = 0.0
    for o in [out] + list(out["aux_outputs"]):
        tot = tot + (o["pred_logits"] * torch.randn(B, Q, 2, generator=g)).sum() + (o["pred_boxes"] * torch.randn(B, Q, 2, generator=g)).sum() \
            + (o["pred_feature"] * torch.randn(B, Q, 2048, generator=g)).sum() * 0.05
    return tot


