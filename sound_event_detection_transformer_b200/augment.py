"""Training-time input transforms of the reference on the device (csrc/augment.cu):

  draw_augment_params / augment_clips   TimeMask -> FreqMask(fill_mode="mean") -> FreqShift of `get_transforms(...,
                                        freq_mask, freq_shift, time_mask)` (utilities/BoxTransforms.py:363-452,471-478).  The
                                        random draws are made on the host from np.random in the reference's order (a seeded run
                                        reproduces the reference's bands); one launch applies them to the whole batch.
  query_patches                         `Query` (utilities/BoxTransforms.py:315-360): SP-SEDT's patch crop + resize to 128 x 64
                                        for every (clip, box), one launch; bit-exact with the reference's PIL path.
  mixup_data / mixup_label_unlabel      utilities/mixup.py:13-127,129-190: the label bookkeeping on the host (it is a handful of
                                        Python decisions per clip), the batch tensor assembled by ONE gather-and-mix launch
                                        instead of ~B slice / unsqueeze / cat ops.

There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib


# ---- TimeMask / FreqMask / FreqShift --------------------------------------------------------------------------------------
def draw_augment_params(n_frames: int, n_mels: int, rng=np.random, time_mask: bool = True, freq_mask: bool = True,
                        freq_shift: bool = True, time_mask_args=(0.0, 0.1, 0.2), freq_mask_args=(0.03, 0.4, 0.5),
                        freq_shift_args=(0.5, 4, 0.0, 2.0), fill_mode: str = "mean", fill_constant: float = 0.0) -> dict:
    """The random draws of ONE clip going through TimeMask(), FreqMask(fill_mode), FreqShift() in the reference's order
    (randomize_parameters at BoxTransforms.py:381-384, :413-416, :440-445), turned into the integer bands the transforms use.
    *_args = (min, max, p) / (min, max, p) / (p, max_band, mean, std), the constructors' defaults."""
    out = dict(tm_t0=0, tm_t=0, fm_f0=0, fm_f=0, fm_mode=0, fm_const=float(fill_constant), fs_shift=0)
    if time_mask:
        apply = rng.uniform(0, 1) < time_mask_args[2]
        t = rng.uniform(time_mask_args[0], time_mask_args[1])
        t0 = rng.uniform(0, 1 - t)
        if apply:
            out["tm_t"], out["tm_t0"] = int(t * n_frames), int(t0 * n_frames)
    if freq_mask:
        apply = rng.uniform(0, 1) < freq_mask_args[2]
        f = rng.uniform(freq_mask_args[0], freq_mask_args[1])
        f0 = rng.uniform(0, 1 - f)
        if apply:
            out["fm_f"], out["fm_f0"] = int(f * n_mels), int(f0 * n_mels)
            out["fm_mode"] = 2 if fill_mode == "mean" else 1
    if freq_shift:
        apply = rng.uniform(0, 1) < freq_shift_args[0]
        s = int(rng.normal(freq_shift_args[2], freq_shift_args[3]))
        while abs(s) > freq_shift_args[1]:
            s = int(rng.normal(freq_shift_args[2], freq_shift_args[3]))
        if apply:
            out["fs_shift"] = s
    return out


def augment_clips(x: torch.Tensor, params: Sequence[dict]) -> torch.Tensor:
    """x: [B, T, F] or [B, 1, T, F] fp32 CUDA, the padded log-mel clips BEFORE Normalize (the reference applies the masks
    between PadOrTrunc and ToTensor).  params: one draw_augment_params() dict per clip.  In place; returns x."""
    if not x.is_cuda:
        raise RuntimeError("augment_clips needs CUDA tensors (there is no CPU path)")
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise RuntimeError("augment_clips expects a contiguous fp32 tensor")
    B, T, F = x.shape[0], x.shape[-2], x.shape[-1]
    if len(params) != B:
        raise ValueError(f"{len(params)} parameter records for {B} clips")
    arr = (_lib.SedtAugmentParams * B)()
    for i, p in enumerate(params):
        arr[i].tm_t0, arr[i].tm_t = p["tm_t0"], p["tm_t"]
        arr[i].fm_f0, arr[i].fm_f, arr[i].fm_mode, arr[i].fm_const = p["fm_f0"], p["fm_f"], p["fm_mode"], p["fm_const"]
        arr[i].fs_shift = p["fs_shift"]
    dev_params = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(x.device)
    scratch = torch.empty(B * T, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.sedt_augment_clips(x.data_ptr(), dev_params.data_ptr(), B, T, F, scratch.data_ptr(), _lib.current_stream()))
    return x


# ---- Query: SP-SEDT patch crop + resize -----------------------------------------------------------------------------------
def patch_bounds(boxes: np.ndarray, n_frames: int, fixed_patch_size: bool = False) -> np.ndarray:
    """(s_idx, e_idx) per (center, width) box with Query.transform_label's float32 arithmetic (BoxTransforms.py:339-350)."""
    b = np.asarray(boxes, np.float32).reshape(-1, 2)
    c, l = b[:, 0], b[:, 1]
    t = np.float32(n_frames)
    s_idx = ((c - l / np.float32(2)) * t).astype(np.int64)          # int(): truncation toward zero
    e_idx = ((c + l / np.float32(2)) * t).astype(np.int64)
    if fixed_patch_size:
        e_idx = np.minimum(n_frames, s_idx + 128)
        s_idx = e_idx - 128
    else:
        empty = s_idx >= e_idx
        s_idx = np.where(empty, np.maximum(0, s_idx - 1), s_idx)
        e_idx = np.where(empty, np.minimum(n_frames, e_idx + 1), e_idx)
    if (s_idx < 0).any() or (e_idx > n_frames).any() or (s_idx >= e_idx).any():
        raise ValueError("a patch box leaves the clip (the reference's Query slices an empty patch and fails there too)")
    return np.stack([s_idx, e_idx], axis=-1).astype(np.int32)


def query_patches(x: torch.Tensor, boxes, fixed_patch_size: bool = False) -> torch.Tensor:
    """x [B, 1, T, F] fp32 CUDA (the normalised clips), boxes [B, P, 2] (center, width) -> patches [B, P, 1, 128, F]:
    label["patches"] of every clip as Query.transform_label builds it, stacked over the batch (engine.py:56-58)."""
    if not x.is_cuda:
        raise RuntimeError("query_patches needs CUDA tensors (there is no CPU path)")
    x = x.to(torch.float32).contiguous()
    B, _, T, F = x.shape
    bx = np.asarray(boxes.detach().cpu() if torch.is_tensor(boxes) else boxes, np.float32).reshape(B, -1, 2)
    P = bx.shape[1]
    bounds = torch.from_numpy(patch_bounds(bx.reshape(-1, 2), T, fixed_patch_size)).to(x.device)
    out = torch.empty(B, P, 1, 128, F, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.sedt_query_patches(x.data_ptr(), bounds.data_ptr(), out.data_ptr(), B, P, T, F, int(bool(fixed_patch_size)),
                                          _lib.current_stream()))
    return out


# ---- mixup ------------------------------------------------------------------------------------------------------------------
def _mix_rows(x: torch.Tensor, rows: List[tuple]) -> torch.Tensor:
    arr = (_lib.SedtMixRow * max(len(rows), 1))()
    for k, (i1, i2, a, b) in enumerate(rows):
        arr[k].i1, arr[k].i2, arr[k].a, arr[k].b = int(i1), int(i2), float(a), float(b)
    table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(x.device)
    out = torch.empty((len(rows),) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
    row_elems = int(x[0].numel())
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.sedt_mix_rows(x.data_ptr(), out.data_ptr(), table.data_ptr(), len(rows), row_elems, _lib.current_stream()))
    return out


def _overlaps_same_class(labels: torch.Tensor, boxes: torch.Tensor) -> bool:
    """mixup.py:83-93: two events of one class overlap in time after the mix."""
    for e in set(labels.tolist()):
        b = boxes[(labels == e)[:len(boxes)]]
        c, l = b.unbind(-1)
        se = torch.stack([c - l / 2, c + l / 2], dim=-1)
        se = se[se.argsort(dim=0)[:, 0]]
        if not (se[:, 1][:-1] < se[:, 0][1:]).all().item():
            return True
    return False


def mixup_data(x, y, mask_strong: slice, mask_weak: Optional[slice], mix_up_ratio: float = 0.5, max_events: int = 20, alpha: float = 3,
               rng=np.random):
    """utilities/mixup.py:13-127 with the same arguments and return value (x, y, strong slice, weak slice).  x: NestedTensor-like
    (`.tensors` [B, C, T, F] on a CUDA device) -- `.tensors` is replaced like the reference does; y: sequence of label dicts."""
    lam = rng.beta(alpha, alpha) if alpha > 0.0 else 1.0
    xt = x.tensors
    if not xt.is_cuda:
        raise RuntimeError("mixup_data needs CUDA tensors (there is no CPU path)")
    bs = xt.shape[0]
    mix_num = int(bs * mix_up_ratio)
    index = np.asarray(list(range(bs)))
    rng.shuffle(index)
    dev = xt.device
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
                weak_l.append({"labels": torch.cat((l1["labels"], l2["labels"]), dim=0), "boxes": torch.tensor([], device=dev),
                               "ratio": torch.tensor([lam] * len(l1["labels"]) + [1 - lam] * len(l2["labels"]), device=dev),
                               "orig_size": l1["orig_size"]})
                weak.append((i, j, lam, 1 - lam))
        elif n1 + n2 > max_events:
            strong_l.append(l1); strong.append((i, i, 1.0, 0.0))
        else:
            lab = {"labels": torch.cat((l1["labels"], l2["labels"]), dim=0), "boxes": torch.cat((l1["boxes"], l2["boxes"]), dim=0),
                   "ratio": torch.tensor([lam] * len(l1["labels"]) + [1 - lam] * len(l2["labels"]), device=dev),
                   "orig_size": l1["orig_size"]}
            if _overlaps_same_class(lab["labels"], lab["boxes"]):
                strong_l.append(l1); strong.append((i, i, 1.0, 0.0))
            else:
                strong_l.append(lab); strong.append((i, j, lam, 1 - lam))
    n_strong = mask_strong.stop
    for i in range(mix_num, n_strong):
        strong_l.append(y[i]); strong.append((i, i, 1.0, 0.0))
    rows, labels = list(strong), list(strong_l)
    if mask_weak is not None:
        ws = mask_weak.stop
        for i in range(n_strong + max(0, mix_num - n_strong), ws):
            weak_l.append(y[i]); weak.append((i, i, 1.0, 0.0))
        for i in range(ws + max(0, mix_num - ws), bs):
            unl_l.append(y[i]); unl.append((i, i, 1.0, 0.0))
        rows += weak + unl
        labels += weak_l + unl_l
    x.tensors = _mix_rows(xt.to(torch.float32).contiguous(), rows)
    return x, labels, slice(len(strong_l)), slice(len(strong_l), len(strong_l) + len(weak_l))


def mixup_label_unlabel(x1, x2, y1, y2, mix_up_ratio: float = 0.5, max_events: int = 20, alpha: float = 3, rng=np.random):
    """utilities/mixup.py:129-190: mixes the first mix_num labelled clips into the (pseudo-labelled) unlabelled batch."""
    assert mix_up_ratio <= 0.5
    lam = rng.beta(alpha, alpha) if alpha > 0.0 else 1.0
    a, b = x1.tensors, x2.tensors
    if not (a.is_cuda and b.is_cuda):
        raise RuntimeError("mixup_label_unlabel needs CUDA tensors (there is no CPU path)")
    bs = a.shape[0]
    mix_num = int(bs * mix_up_ratio)
    dev = a.device
    n1_rows = a.shape[0]                              # rows of the concatenated [x1; x2] source: x2 rows are offset by n1_rows
    rows, labels = [], []
    for i in range(mix_num):
        l1, l2 = y1[i], y2[i]
        if len(l1["boxes"]) + len(l2["boxes"]) > max_events:
            if len(l2["boxes"]):
                labels.append(l2); rows.append((n1_rows + i, n1_rows + i, 1.0, 0.0))
            else:
                labels.append(l1); rows.append((i, i, 1.0, 0.0))
            continue
        lab = {"labels": torch.cat((l1["labels"], l2["labels"]), dim=0), "boxes": torch.cat((l1["boxes"], l2["boxes"]), dim=0),
               "ratio": torch.tensor([lam] * len(l1["labels"]) + [1 - lam] * len(l2["labels"]), device=dev),
               "orig_size": l1["orig_size"]}
        if _overlaps_same_class(lab["labels"], lab["boxes"]):
            labels.append(l1); rows.append((i, i, 1.0, 0.0))
        else:
            labels.append(lab); rows.append((i, n1_rows + i, lam, 1 - lam))
    for i in range(mix_num, b.shape[0]):
        labels.append(y2[i]); rows.append((n1_rows + i, n1_rows + i, 1.0, 0.0))
    src = torch.cat([a.to(torch.float32), b.to(torch.float32)], dim=0).contiguous()
    x2.tensors = _mix_rows(src, rows)
    return x2, labels
