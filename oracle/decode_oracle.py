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
