"""The evaluation input pipeline of the reference on the device: `get_transforms(frames, scaler, add_axis=0)`
(utilities/BoxTransforms.py:454-490) = ApplyLog -> PadOrTrunc -> ToTensor -> Normalize, one launch for the whole batch
(csrc/prepare.cu).  The random training augmentations (TimeMask, FreqMask, FreqShift, mixup) and SP-SEDT's patch crop +
resize are not covered."""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib


def prepare_clips(raw: Sequence[torch.Tensor], frames: int, mean: Optional[torch.Tensor] = None,
                  std: Optional[torch.Tensor] = None, apply_log: bool = True, device: Optional[torch.device] = None) -> torch.Tensor:
    """raw: per-clip mel amplitude features [T_i, F] (what SedData stores); mean / std: Scaler.mean_ / Scaler.std_ ([F], float64).
    Returns the model input [B, 1, frames, F] fp32 on the device."""
    lib = _lib.load()
    dev = torch.device(device) if device is not None else (raw[0].device if raw[0].is_cuda else torch.device("cuda"))
    if dev.type != "cuda":
        raise RuntimeError("prepare_clips needs a CUDA device (there is no CPU path)")
    F = int(raw[0].shape[-1])
    lens = [int(r.shape[0]) for r in raw]
    cat = torch.cat([r.reshape(-1, F).to(torch.float32) for r in raw]).to(dev, non_blocking=True).contiguous()
    off = [0]
    for n in lens:
        off.append(off[-1] + n)
    offsets = torch.tensor(off, dtype=torch.int64).to(dev, non_blocking=True)
    m = s = None
    if mean is not None:
        m = torch.as_tensor(mean, dtype=torch.float64).reshape(-1).to(dev).contiguous()
        s = torch.as_tensor(std, dtype=torch.float64).reshape(-1).to(dev).contiguous()
        if m.numel() != F or s.numel() != F:
            raise ValueError(f"scaler mean / std must have {F} entries")
    out = torch.empty(len(raw), 1, frames, F, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sedt_prepare_clips(cat.data_ptr(), offsets.data_ptr(), _lib.ptr(m) or None, _lib.ptr(s) or None, out.data_ptr(),
                                          len(raw), int(frames), F, int(bool(apply_log)), _lib.current_stream()))
    return out
