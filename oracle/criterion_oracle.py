"""TEST INFRASTRUCTURE ONLY -- numpy fp32 restatement of SetCriterion.forward for the supervised default recipe
(sedt/sedt.py:309-352 with fine_tune = normalize = fl = False): matcher per decoder layer (oracle/matcher_oracle.py),
loss_labels (:188-221), loss_boxes (:238-261), loss_cardinality (:223-236), loss_weak (:161-186).
Pinned against the reference's own SetCriterion by tests/test_oracle_golden.py on the fixtures criterion_*.npz
(tests/golden/make_golden.py: run_criterion).  Only tests/ may import this module."""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np

from . import matcher_oracle

f32 = np.float32


def _log_softmax(x: np.ndarray) -> np.ndarray:
    x = x.astype(f32)
    z = x - x.max(-1, keepdims=True)
    return (z - np.log(np.exp(z, dtype=f32).sum(-1, keepdims=True, dtype=f32))).astype(f32)


def _layer(logits, boxes, targets, num_classes, eos_coef, num_boxes, log):
    B, Q, C1 = logits.shape
    idx, coef = matcher_oracle.hungarian_matcher({"pred_logits": logits, "pred_boxes": boxes}, targets)
    out = {}
    # loss_labels
    cls = np.full((B, Q), num_classes, np.int64)
    for b, (r, c) in enumerate(idx):
        cls[b, r] = np.asarray(targets[b]["labels"])[c]
    w = np.ones(C1, f32); w[-1] = f32(eos_coef)
    logp = _log_softmax(logits)
    ce = -w[cls] * np.take_along_axis(logp, cls[..., None], -1)[..., 0]
    out["loss_ce"] = f32(ce.sum(dtype=f32) / f32(num_boxes))
    if log:
        pairs = [(b, q, np.asarray(targets[b]["labels"])[c_]) for b, (r, c) in enumerate(idx) for q, c_ in zip(r, c)]
        if pairs:
            acc = np.mean([logits[b, q].argmax() == t for b, q, t in pairs]) * 100.0
            out["class_error"] = f32(100.0 - acc)
        else:
            out["class_error"] = f32(100.0)
    # loss_cardinality
    n_pred = (logits.argmax(-1) != C1 - 1).sum(1).astype(f32)
    n_tgt = np.asarray([len(t["labels"]) for t in targets], f32)
    out["cardinality_error"] = f32(np.abs(n_pred - n_tgt).mean())
    # loss_boxes on (s, 0, e, 1)
    l1 = f32(0); gi = f32(0)
    for b, (r, c) in enumerate(idx):
        if len(r) == 0:
            continue
        p = boxes[b][r].astype(f32); t = np.asarray(targets[b]["boxes"], f32).reshape(-1, 2)[c]
        s1, e1 = p[:, 0] - p[:, 1] / f32(2), p[:, 0] + p[:, 1] / f32(2)
        s2, e2 = t[:, 0] - t[:, 1] / f32(2), t[:, 0] + t[:, 1] / f32(2)
        l1 += (np.abs(s1 - s2) + np.abs(e1 - e2)).sum(dtype=f32)
        inter = np.maximum(np.minimum(e1, e2) - np.maximum(s1, s2), f32(0))
        union = (e1 - s1) + (e2 - s2) - inter
        enc = np.maximum(np.maximum(e1, e2) - np.minimum(s1, s2), f32(0))
        gi += (f32(1) - (inter / union - (enc - union) / enc)).sum(dtype=f32)
    out["loss_bbox"] = f32(l1 / f32(num_boxes))
    out["loss_giou"] = f32(gi / f32(num_boxes))
    return out, idx


def set_criterion(outputs: Dict, targets: Sequence[dict], num_classes: int, eos_coef: float) -> Dict[str, np.float32]:
    """outputs: pred_logits [B,Q,C+1], pred_boxes [B,Q,2], optional at [B,C], aux_outputs (list of dicts); every clip strong."""
    lg = np.asarray(outputs["pred_logits"], f32)
    Q = lg.shape[1]
    num_boxes = float(sum(min(Q, len(t["boxes"])) for t in targets))
    losses, _ = _layer(lg, np.asarray(outputs["pred_boxes"], f32), targets, num_classes, eos_coef, num_boxes, True)
    if "at" in outputs:
        p = np.asarray(outputs["at"], f32).reshape(len(targets), -1)
        gt = np.zeros_like(p)
        for b, t in enumerate(targets):
            for l in np.asarray(t["labels"]).reshape(-1):
                gt[b, int(l)] = 1.0
        bce = -(gt * np.maximum(np.log(p), f32(-100)) + (1 - gt) * np.maximum(np.log(f32(1) - p), f32(-100)))
        losses["loss_weak"] = f32(bce.mean(dtype=f32))
    for i, aux in enumerate(outputs.get("aux_outputs", [])):
        part, _ = _layer(np.asarray(aux["pred_logits"], f32), np.asarray(aux["pred_boxes"], f32), targets, num_classes, eos_coef,
                         num_boxes, False)
        losses.update({f"{k}_{i}": v for k, v in part.items()})
    return losses
