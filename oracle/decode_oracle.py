"""ORACLE (test infrastructure, not product code).

numpy restatement of BoxEncoder.decode_strong (utilities/BoxEncoder.py:179-226),
the consumer the north_star names for the "decoded events identical" check:
keep queries with score >= threshold and duration >= 0.2 s, then per class
sort by onset and drop the lower-scored one of each overlapping pair.
Pinned against the reference's BoxEncoder by tests/golden/make_golden.py
(fixture `events_*.json`).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def decode_strong(result: dict, class_names: Sequence[str], threshold: float = 0.5, del_overlap: bool = True) -> List[list]:
    scores = np.asarray(result["scores"]); labels = np.asarray(result["labels"]); boxes = np.asarray(result["boxes"])
    out: List[list] = []
    nq = len(scores)
    if not del_overlap:                                            # BoxEncoder.py:192-199
        for i in range(nq):
            if scores[i] > threshold:
                onset, offset = boxes[i]
                if offset - onset >= 0.2:
                    out.append([class_names[labels[i]], onset, offset, scores[i]])
        return out
    ev = {}
    for i in range(nq):                                            # BoxEncoder.py:202-209
        if scores[i] >= threshold:
            onset, offset = boxes[i]
            if offset - onset >= 0.2:
                ev.setdefault(class_names[labels[i]], []).append(np.asarray([scores[i], onset, offset]))
    for name in ev:                                                # BoxEncoder.py:212-225
        arr = np.vstack(ev[name])
        arr = arr[np.argsort(arr, axis=0)[:, 1]]
        i = 1
        while i < len(arr):
            if arr[i][1] < arr[i - 1][2]:
                arr = np.delete(arr, i - 1 if arr[i][0] > arr[i - 1][0] else i, axis=0)
                continue
            i += 1
        for r in arr:
            out.append([name, r[1], r[2], r[0]])
    return out


def pseudo_labels(logits: np.ndarray, boxes: np.ndarray, at, class_thr: np.ndarray, clip_seconds: float,
                  del_overlap: bool = True):
    """numpy fp32 restatement of engine.get_pseudo_labels (engine.py:300-348) for one batch: PostProcess with at_m = 1 and
    is_semi (sedt/sedt.py:359-396), class-wise threshold + minimum width, greedy same-class overlap suppression in
    descending score order.  Returns per clip (labels int64 [n], boxes fp32 [n, 2] (center, width)).  Pinned against the
    reference's own function by tests/golden/make_golden.py (fixture `pseudo_*.npz`)."""
    f32 = np.float32
    logits = np.asarray(logits, f32); boxes = np.asarray(boxes, f32); class_thr = np.asarray(class_thr, f32)
    x = logits - logits.max(-1, keepdims=True)
    e = np.exp(x, dtype=f32)
    prob = (e / e.sum(-1, keepdims=True, dtype=f32)).astype(f32)
    ev = prob[..., :-1]
    if at is not None:
        tags = (np.asarray(at, f32) >= class_thr).astype(f32)
        ev = ev * tags[:, None, :]
    out = []
    min_w = f32(0.2 / clip_seconds)
    for b in range(logits.shape[0]):
        scores = ev[b].max(-1); labels = ev[b].argmax(-1)
        keep0 = (scores >= class_thr[labels]) & (boxes[b, :, 1] > min_w)
        lab, bx, sc = labels[keep0], boxes[b][keep0], scores[keep0]
        if not del_overlap:
            out.append((lab.astype(np.int64), bx)); continue
        order = list(np.argsort(-sc, kind="stable"))
        xs = bx[:, 0] - bx[:, 1] / f32(2); ys = bx[:, 0] + bx[:, 1] / f32(2)
        keep = []
        while order:
            k = order[0]; keep.append(k)
            order = [j for j in order[1:] if max(f32(0), min(ys[j], ys[k]) - max(xs[j], xs[k])) == 0 or lab[j] != lab[k]]
        out.append((lab[keep].astype(np.int64), bx[keep]))
    return out
