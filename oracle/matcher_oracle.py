"""ORACLE (test infrastructure, not product code).

numpy fp32 restatement of HungarianMatcher.forward (sedt/matcher.py:41-133:
default path, focal class cost `fl` :77-82, fine_tune relaxation :99-121,
normalize / ratio coefficients :123-133; box maths from
utilities/box_ops.py:9-14,29-56).  The reference builds one
[B*Q, sum K] matrix and slices the block diagonal (matcher.py:91-95); this
file computes the per-clip blocks directly, which is the same arithmetic on
the entries that are actually used.

The per-clip solve is scipy.optimize.linear_sum_assignment, the reference's
own call (matcher.py:95); oracle/lsap_oracle.c restates the same published
algorithm in C and is pinned against scipy in tests/test_matcher_oracle.py.

Parity pinning: tests/golden/make_golden.py runs the reference matcher on
seeded inputs and stores indices; tests/test_oracle_golden.py compares.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this file.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from collections import Counter
from typing import List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32 = np.float32


def softmax_f32(logits: np.ndarray) -> np.ndarray:
    """matcher.py:65 (`.softmax(-1)` in fp32)."""
    x = logits.astype(_f32)
    x = x - x.max(axis=-1, keepdims=True)
    e = np.exp(x, dtype=_f32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=_f32)).astype(_f32)


ALPHA_FL, GAMMA_FL = 0.5, 1.0        # config.py:71-72


def sigmoid_f32(logits: np.ndarray) -> np.ndarray:
    x = logits.astype(_f32)
    return (_f32(1) / (_f32(1) + np.exp(-x, dtype=_f32))).astype(_f32)


def focal_class_cost(prob: np.ndarray, tgt_labels: np.ndarray) -> np.ndarray:
    """matcher.py:78-82 on sigmoid probabilities."""
    p = prob.astype(_f32)
    neg = (_f32(1 - ALPHA_FL) * p ** _f32(GAMMA_FL)) * (-np.log(_f32(1) - p + _f32(1e-8)))
    pos = (_f32(ALPHA_FL) * (_f32(1) - p) ** _f32(GAMMA_FL)) * (-np.log(p + _f32(1e-8)))
    return (pos[:, tgt_labels] - neg[:, tgt_labels]).astype(_f32)


def cost_block(prob: np.ndarray, boxes: np.ndarray, tgt_labels: np.ndarray, tgt_boxes: np.ndarray,
               w_class: float = 1.0, w_bbox: float = 5.0, w_giou: float = 2.0, fl: bool = False,
               location_only: bool = False) -> np.ndarray:
    """One clip's [Q,K] fp32 cost block.  prob [Q,C+1], boxes [Q,2] (c,l),
    tgt_labels [K], tgt_boxes [K,2].  matcher.py:76,85,88,91; box_ops.py:9-14,29-56."""
    prob = prob.astype(_f32); boxes = boxes.astype(_f32); tgt_boxes = tgt_boxes.astype(_f32)
    half = _f32(2.0)
    sp = (boxes[:, 0] - boxes[:, 1] / half)[:, None]; ep = (boxes[:, 0] + boxes[:, 1] / half)[:, None]
    st = (tgt_boxes[:, 0] - tgt_boxes[:, 1] / half)[None, :]; et = (tgt_boxes[:, 0] + tgt_boxes[:, 1] / half)[None, :]
    cost_class = focal_class_cost(prob, tgt_labels) if fl else -prob[:, tgt_labels]
    cost_bbox = np.abs(sp - st) + np.abs(ep - et)                 # cdist p=1 on (s,0,e,1)
    area_p = ep - sp; area_t = et - st                            # box_area with y-extent 1
    inter = np.maximum(np.minimum(ep, et) - np.maximum(sp, st), _f32(0))
    union = area_p + area_t - inter
    iou = inter / union
    enc = np.maximum(np.maximum(ep, et) - np.minimum(sp, st), _f32(0))
    giou = iou - (enc - union) / enc
    if location_only:                                             # C_l of the fine_tune branch, matcher.py:104
        return (_f32(w_bbox) * cost_bbox + _f32(w_giou) * (-giou)).astype(_f32)
    C = _f32(w_bbox) * cost_bbox + _f32(w_class) * cost_class + _f32(w_giou) * (-giou)
    return C.astype(_f32)


def hungarian_matcher(outputs: dict, targets: Sequence[dict], normalize: bool = False,
                      w_class: float = 1.0, w_bbox: float = 5.0, w_giou: float = 2.0,
                      solver: str = "scipy", fl: bool = False, fine_tune: bool = False, epsilon: float = 1.0,
                      alpha: float = 1.0, rand=None) -> Tuple[List[Tuple[np.ndarray, np.ndarray]], List[np.ndarray]]:
    """Returns (indices, Coef) like matcher.py:97,123-133 but as numpy arrays.  fine_tune: `rand(n)` must return the n
    uniform numbers the reference draws with torch.rand(n) for the clip (matcher.py:116); pass a wrapper of the seeded
    torch generator to reproduce a reference run."""
    logits = np.asarray(outputs["pred_logits"], dtype=_f32)
    boxes = np.asarray(outputs["pred_boxes"], dtype=_f32)
    prob = sigmoid_f32(logits) if fl else softmax_f32(logits)
    Q = logits.shape[1]
    idx, coef = [], []
    for b, tgt in enumerate(targets):
        tb = np.asarray(tgt["boxes"], dtype=_f32).reshape(-1, 2)
        tl = np.asarray(tgt["labels"], dtype=np.int64)[: len(tb)]
        C = cost_block(prob[b], boxes[b], tl, tb, w_class, w_bbox, w_giou, fl=fl)
        r, c = lsap(C, solver)
        r, c = r.astype(np.int64), c.astype(np.int64)
        if fine_tune:                                             # matcher.py:99-121
            Cl = cost_block(prob[b], boxes[b], tl, tb, w_class, w_bbox, w_giou, location_only=True)
            lmin, larg = Cl.min(-1), Cl.argmin(-1)
            num_gt = len(c)
            reserved = lmin < _f32(epsilon)
            keep = reserved[r]
            r, c = r[keep], c[keep]
            reserved[r] = False
            cand = np.where(reserved)[0]
            drop = np.asarray(rand(len(cand)), dtype=_f32) > (alpha * num_gt / Q)
            reserved[cand[drop]] = False
            r = np.concatenate([r, np.arange(Q)[reserved]]); c = np.concatenate([c, larg[reserved]])
        idx.append((r, c))
        if normalize:
            cnt = Counter(c.tolist())
            coef.append(np.array([1.0 / cnt[j] for j in c.tolist()], dtype=_f32))
        elif "ratio" in tgt:
            coef.append(np.asarray(tgt["ratio"], dtype=_f32))
        else:
            coef.append(np.ones(len(c), dtype=_f32))
    return idx, coef


def lsap(C: np.ndarray, solver: str = "scipy"):
    if solver == "scipy":
        from scipy.optimize import linear_sum_assignment
        return linear_sum_assignment(C)
    return lsap_c(C)


# ---- the C restatement ----------------------------------------------------
_lib = None


def build_c(force: bool = False) -> str:
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liblsap_oracle.so")
    src = os.path.join(_HERE, "lsap_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src, "-lm"])
    return so


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c())
        _lib.lsap_oracle_f64.restype = ctypes.c_int
        _lib.lsap_oracle_batched_f32.restype = ctypes.c_int
    return _lib


def lsap_c(C: np.ndarray):
    lib = _load()
    C64 = np.ascontiguousarray(C, dtype=np.float64)
    nr, nc = C64.shape
    n = min(nr, nc)
    a = np.zeros(n, dtype=np.int64); b = np.zeros(n, dtype=np.int64)
    rc = lib.lsap_oracle_f64(ctypes.c_int64(nr), ctypes.c_int64(nc), C64.ctypes.data_as(ctypes.c_void_p),
                             a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError("cost matrix is infeasible" if rc == -1 else "matrix contains invalid numeric entries")
    return a, b


def lsap_c_batched(cost: np.ndarray, K: np.ndarray):
    """cost [B,Q,ldk] fp32, K [B] int32 -> rows, cols [B,min(Q,ldk)] int64, counts [B] int32."""
    lib = _load()
    cost = np.ascontiguousarray(cost, dtype=_f32)
    K = np.ascontiguousarray(K, dtype=np.int32)
    B, Q, ldk = cost.shape
    cap = min(Q, ldk)
    rows = np.full((B, cap), -1, dtype=np.int64); cols = np.full((B, cap), -1, dtype=np.int64)
    counts = np.zeros(B, dtype=np.int32)
    rc = lib.lsap_oracle_batched_f32(ctypes.c_int64(B), ctypes.c_int64(Q), ctypes.c_int64(ldk),
                                     cost.ctypes.data_as(ctypes.c_void_p), K.ctypes.data_as(ctypes.c_void_p),
                                     rows.ctypes.data_as(ctypes.c_void_p), cols.ctypes.data_as(ctypes.c_void_p),
                                     counts.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        raise ValueError(f"lsap oracle failed rc={rc}")
    return rows, cols, counts
