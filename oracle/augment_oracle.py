"""ORACLE (test infrastructure, not product code).

numpy restatement of the training-time input transforms of the reference:

  time_mask / freq_mask / freq_shift   utilities/BoxTransforms.py:363-452 (TimeMask, FreqMask(fill_mode="mean"|"constant"),
                                       FreqShift), including the ORDER of their np.random draws (draw_params), so a seeded
                                       run reproduces the reference's parameters;
  query_patches                        utilities/BoxTransforms.py:315-360 (Query: crop by (center, width) box, min/max
                                       normalise, ToPILImage -> Resize((128, 64)) -> ToTensor, de-normalise);
  pil_resize_rows                      the arithmetic behind transforms.Resize on a mode-"L" image, restated from Pillow's
                                       published algorithm (src/libImaging/Resample.c: precompute_coeffs with the bilinear
                                       filter, support scaled by the down-sampling factor = antialiasing; 8-bit fixed point
                                       with PRECISION_BITS = 22; vertical pass only because the width stays 64).  Pillow is a
                                       third-party dependency the reference does not pin (12.2.0 in this image);
  mixup_plan / mixup_data              utilities/mixup.py:13-127 (mixup_data): the label bookkeeping and
                                       lam * x1 + (1 - lam) * x2 in fp32 (two rounded products, one rounded sum).

Pinned against the reference's own classes by tests/golden/make_golden.py (fixtures augment_*.npz, query_*.npz, mixup_*.npz)
and tests/test_oracle_golden.py.  Only tests/ may import this module.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32


# ---- TimeMask / FreqMask / FreqShift ------------------------------------------------------------------------------
def draw_params(rng=np.random, time_mask=True, freq_mask=True, freq_shift=True, tm=(0.0, 0.1, 0.2), fm=(0.03, 0.4, 0.5),
                fs=(0.5, 4, 0.0, 2.0)) -> dict:
    """The draws of one sample going through Compose([... TimeMask(), FreqMask(fill_mode="mean"), FreqShift() ...]) in the
    reference's order (randomize_parameters of each transform: :381-384, :413-416, :440-445)."""
    p = {}
    if time_mask:
        p["tm_apply"] = rng.uniform(0, 1) < tm[2]
        p["tm_t"] = rng.uniform(tm[0], tm[1])
        p["tm_t0"] = rng.uniform(0, 1 - p["tm_t"])
    if freq_mask:
        p["fm_apply"] = rng.uniform(0, 1) < fm[2]
        p["fm_f"] = rng.uniform(fm[0], fm[1])
        p["fm_f0"] = rng.uniform(0, 1 - p["fm_f"])
    if freq_shift:
        p["fs_apply"] = rng.uniform(0, 1) < fs[0]
        s = int(rng.normal(fs[2], fs[3]))
        while abs(s) > fs[1]:
            s = int(rng.normal(fs[2], fs[3]))
        p["fs_shift"] = s
    return p


def time_mask(data: np.ndarray, t_frac: float, t0_frac: float, fade: bool = False) -> np.ndarray:
    data = data.copy()
    n = data.shape[0]
    t, t0 = int(t_frac * n), int(t0_frac * n)
    mask = np.zeros((t, data.shape[1]))
    if fade:
        fl = int(t * 0.1)
        mask[0:fl, :] = np.linspace(1, 0, num=fl)[:, None] if fl else mask[0:fl, :]
        if fl:
            mask[-fl:, :] = np.linspace(0, 1, num=fl)[:, None]
    data[t0:t0 + t, :] *= mask.astype(data.dtype) if not fade else mask
    return data


def freq_mask(data: np.ndarray, f_frac: float, f0_frac: float, fill_mode: str = "mean", constant: float = 0.0) -> np.ndarray:
    data = data.copy()
    nmel = data.shape[1]
    f, f0 = int(f_frac * nmel), int(f0_frac * nmel)
    fill = np.mean(data[:, f0:f0 + f]) if fill_mode == "mean" else constant
    data[:, f0:f + f0] = fill
    return data


def _pairwise_sum_f32(a: np.ndarray) -> np.float32:
    """numpy/core/src/umath/loops_utils.h.src: FLOAT_pairwise_sum."""
    n = len(a)
    if n < 8:
        res = f32(0)
        for v in a:
            res = f32(res + v)
        return res
    if n <= 128:
        r = [f32(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = f32(r[j] + a[i + j])
            i += 8
        res = f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])))
        while i < n:
            res = f32(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return f32(_pairwise_sum_f32(a[:n2]) + _pairwise_sum_f32(a[n2:]))


def numpy_mean_f32(s: np.ndarray) -> np.float32:
    """np.mean of a non-contiguous [T, n] float32 slice, operation by operation (what csrc/augment.cu reproduces): C-order
    chunks of (8192 // n) * n elements through the pairwise sum, chunk sums accumulated in float32, float32(double(sum) / count).
    Checked against np.mean itself in tests/test_oracle_golden.py."""
    n = s.shape[1]
    chunk = (8192 // n) * n
    flat = np.ascontiguousarray(s).reshape(-1)
    acc = f32(0)
    for i in range(0, len(flat), chunk):
        acc = f32(acc + _pairwise_sum_f32(flat[i:i + chunk]))
    return f32(np.float64(acc) / np.float64(s.size))


def freq_shift(data: np.ndarray, shift: int) -> np.ndarray:
    data = np.roll(data, shift, axis=1)
    if shift >= 0:
        data[:, :shift] = 0
    else:
        data[:, shift:] = 0
    return data


def augment(data: np.ndarray, p: dict) -> np.ndarray:
    """[T, F] fp32 log-mel clip (after PadOrTrunc) through the three transforms with the drawn parameters."""
    if p.get("tm_apply"):
        data = time_mask(data, p["tm_t"], p["tm_t0"])
    if p.get("fm_apply"):
        data = freq_mask(data, p["fm_f"], p["fm_f0"], p.get("fm_mode", "mean"), p.get("fm_const", 0.0))
    if p.get("fs_apply"):
        data = freq_shift(data, p["fs_shift"])
    return data


# ---- Pillow's resample (vertical pass, 8 bits per channel, bilinear) ------------------------------------------------
PRECISION_BITS = 32 - 8 - 2


def _bilinear(x: float) -> float:
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def resample_coeffs(in_size: int, out_size: int) -> Tuple[List[int], List[List[int]]]:
    """precompute_coeffs + normalize_coeffs_8bpc: per output index the first input index and the integer weights."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ss = 1.0 / filterscale
    firsts, weights = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [_bilinear((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(k)
        k = [w / ww if ww != 0.0 else w for w in k]
        ik = [int(w * (1 << PRECISION_BITS) - 0.5) if w < 0 else int(w * (1 << PRECISION_BITS) + 0.5) for w in k]
        firsts.append(xmin)
        weights.append(ik)
    return firsts, weights


def pil_resize_rows(img: np.ndarray, out_rows: int) -> np.ndarray:
    """img uint8 [h, w] -> uint8 [out_rows, w] like Image.resize((w, out_rows), BILINEAR) (ImagingResampleVertical_8bpc)."""
    h, w = img.shape
    if h == out_rows:
        return img.copy()
    firsts, weights = resample_coeffs(h, out_rows)
    out = np.empty((out_rows, w), np.uint8)
    src = img.astype(np.int64)
    for yy in range(out_rows):
        acc = np.full(w, 1 << (PRECISION_BITS - 1), np.int64)
        for j, kw in enumerate(weights[yy]):
            acc += src[firsts[yy] + j] * kw
        out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def patch_bounds(c: np.float32, l: np.float32, t: int, fixed: bool) -> Tuple[int, int]:
    """Query.transform_label :339-350 with its float32 arithmetic (box.numpy() yields float32 scalars)."""
    c, l = f32(c), f32(l)
    s, e = c - l / f32(2), c + l / f32(2)
    s_idx, e_idx = int(s * f32(t)), int(e * f32(t))
    if fixed:
        e_idx = min(t, s_idx + 128)
        s_idx = e_idx - 128
    elif s_idx >= e_idx:
        s_idx = max(0, s_idx - 1)
        e_idx = min(t, e_idx + 1)
    return s_idx, e_idx


def query_patches(data: np.ndarray, boxes: np.ndarray, fixed_patch_size: bool = False) -> np.ndarray:
    """data [1, T, F] fp32 (normalised clip), boxes [P, 2] fp32 (center, width) -> patches [P, 1, 128, F] fp32."""
    _, t, F = data.shape
    out = []
    for c, l in np.asarray(boxes, f32):
        s_idx, e_idx = patch_bounds(c, l, t, fixed_patch_size)
        if fixed_patch_size:
            out.append(data[:, s_idx:e_idx, :].astype(f32))
            continue
        ori = data[0, s_idx:e_idx, :].astype(f32)
        mn, mx = ori.min(), ori.max()
        norm = (ori - mn) / (mx - mn)
        u8 = (norm * f32(255)).astype(np.uint8)               # ToPILImage: pic.mul(255).byte()
        res = pil_resize_rows(u8, 128)
        back = res.astype(f32) / f32(255)                     # ToTensor: byte -> float, div(255)
        out.append((back * (mx - mn) + mn)[None].astype(f32))
    return np.stack(out)


# ---- mixup (utilities/mixup.py:13-127) -------------------------------------------------------------------------------
def _se(boxes: np.ndarray) -> np.ndarray:
    c, l = boxes[:, 0], boxes[:, 1]
    return np.stack([c - f32(0.5) * l, c + f32(0.5) * l], axis=-1)


def _same_class_overlap(labels: np.ndarray, boxes: np.ndarray) -> bool:
    for e in set(labels.tolist()):
        b = _se(boxes[(labels == e)[:len(boxes)]])
        b = b[np.argsort(b[:, 0], kind="stable")]
        if not (b[:, 1][:-1] < b[:, 0][1:]).all():
            return True
    return False


def mixup_plan(y: Sequence[dict], n_strong: int, n_weak: Optional[int], lam: float, index: np.ndarray, mix_up_ratio: float = 0.5,
               max_events: int = 20):
    """The label bookkeeping of mixup_data for a batch laid out [strong | weak | unlabelled] (mask_strong = slice(n_strong),
    mask_weak = slice(n_strong, n_strong + n_weak) or None).  Returns (rows, labels, n_strong_out, n_weak_out) where every
    output row is (i1, i2, a, b): out = a * x[i1] + b * x[i2] (b = 0, i2 = i1 for an unmixed row)."""
    bs = len(y)
    mix_num = int(bs * mix_up_ratio)
    strong, weak, unl = [], [], []
    strong_l, weak_l, unl_l = [], [], []
    for i in range(mix_num):
        j = int(index[i])
        l1, l2 = y[i], y[j]
        n1, n2 = len(l1["boxes"]), len(l2["boxes"])
        if n1 == 0 or n2 == 0:
            if n1 > 0:
                strong_l.append(l1); strong.append((i, i, 1.0, 0.0))
            elif n2 > 0:
                strong_l.append(l2); strong.append((j, j, 1.0, 0.0))
            else:
                weak_l.append({"labels": np.concatenate([l1["labels"], l2["labels"]]), "boxes": np.zeros((0,), f32),
                               "ratio": np.asarray([lam] * len(l1["labels"]) + [1 - lam] * len(l2["labels"]), f32),
                               "orig_size": l1["orig_size"]})
                weak.append((i, j, lam, 1 - lam))
        elif n1 + n2 > max_events:
            strong_l.append(l1); strong.append((i, i, 1.0, 0.0))
        else:
            lab = {"labels": np.concatenate([l1["labels"], l2["labels"]]), "boxes": np.concatenate([l1["boxes"], l2["boxes"]]),
                   "ratio": np.asarray([lam] * len(l1["labels"]) + [1 - lam] * len(l2["labels"]), f32), "orig_size": l1["orig_size"]}
            if _same_class_overlap(lab["labels"], lab["boxes"]):
                strong_l.append(l1); strong.append((i, i, 1.0, 0.0))
            else:
                strong_l.append(lab); strong.append((i, j, lam, 1 - lam))
    for i in range(mix_num, n_strong):
        strong_l.append(y[i]); strong.append((i, i, 1.0, 0.0))
    if n_weak is not None:
        ws = n_strong + n_weak
        for i in range(n_strong + max(0, mix_num - n_strong), ws):
            weak_l.append(y[i]); weak.append((i, i, 1.0, 0.0))
        for i in range(ws + max(0, mix_num - ws), bs):
            unl_l.append(y[i]); unl.append((i, i, 1.0, 0.0))
    rows = strong + (weak + unl if n_weak is not None else [])
    labels = strong_l + (weak_l + unl_l if n_weak is not None else [])
    return rows, labels, len(strong_l), len(weak_l)


def mixup_rows(x: np.ndarray, rows) -> np.ndarray:
    out = np.empty((len(rows),) + x.shape[1:], f32)
    for k, (i1, i2, a, b) in enumerate(rows):
        out[k] = x[i1] if b == 0.0 else f32(a) * x[i1] + f32(b) * x[i2]
    return out
