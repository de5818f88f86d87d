"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's evaluation input pipeline
`get_transforms(frames, scaler, add_axis=0)` (utilities/BoxTransforms.py:454-490): ApplyLog -> PadOrTrunc -> ToTensor ->
Normalize.

* pad_trunc / normalize restate utilities/BoxTransforms.py:70-88 and utilities/Scaler.py:102-108 and are PINNED against the
  reference's own PadOrTrunc / ToTensor / Normalize classes (tests/golden/make_golden.py: run_prepare, fixture prepare_*.npz).
* amplitude_to_db restates librosa.amplitude_to_db (called at BoxTransforms.py:67).  librosa is a third-party dependency the
  reference does not pin and that is absent from this image: the function is restated from its published definition
  (librosa.core.spectrum: power_to_db(|S|^2, ref=1, amin=1e-10, top_db=80)) -- PARITY UNPINNED for this one step.
Only tests/ may import this module."""
from __future__ import annotations

import numpy as np


def amplitude_to_db(S: np.ndarray, amin: float = 1e-5, top_db: float = 80.0) -> np.ndarray:
    mag = np.abs(np.asarray(S, np.float32))
    power = mag * mag
    log_spec = (np.float32(10.0) * np.log10(np.maximum(np.float32(amin * amin), power))).astype(np.float32)
    return np.maximum(log_spec, log_spec.max() - np.float32(top_db))


def pad_trunc(x: np.ndarray, frames: int) -> np.ndarray:
    if x.shape[-2] <= frames:
        return np.pad(x, ((0, frames - x.shape[-2]), (0, 0)), mode="constant")
    return x[:frames]


def normalize(x: np.ndarray, mean: np.ndarray, std: np.ndarray) -> np.ndarray:
    return ((x.astype(np.float32) - mean.astype(np.float64)) / std.astype(np.float64)).astype(np.float32)


def prepare_clip(raw: np.ndarray, frames: int, mean=None, std=None, apply_log: bool = True) -> np.ndarray:
    x = amplitude_to_db(raw) if apply_log else np.asarray(raw, np.float32)
    x = pad_trunc(x, frames)[None].astype(np.float32)            # ToTensor(unsqueeze_axis=0).float()
    return normalize(x, mean, std) if mean is not None else x
