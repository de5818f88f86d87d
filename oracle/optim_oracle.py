"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy fp32) of the optimizer half of the reference's training step:
`torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)` followed by `optimizer.step()` with
`torch.optim.AdamW` (engine.py:76-80; two lr groups, weight decay 1e-4: train_sedt.py:234-240,269-270).

The arithmetic lives in third-party code (torch, unpinned by the reference; 2.11.0 here):
torch/nn/utils/clip_grad.py (`clip_coef = max_norm / (total_norm + 1e-6)`, clamped to 1) and
torch/optim/adamw.py `_single_tensor_adamw` (decoupled decay, lerp, addcmul, sqrt / bias_correction2_sqrt + eps,
addcdiv).  Pinned against torch itself on the CPU in tests/test_optim_oracle.py.  Only tests/, smoke() and bench.py's
CPU baseline may import this module; the product never does."""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np

f32 = np.float32


def grad_norm(grads: Sequence[np.ndarray]) -> np.float32:
    """|| [ ||g_i||_2 ] ||_2 in fp32 with fp64 accumulation inside each norm (what at::norm does on CPU)."""
    norms = np.array([np.sqrt(np.sum(g.astype(np.float64) ** 2)) for g in grads], dtype=np.float32)
    return f32(np.sqrt(np.sum(norms.astype(np.float64) ** 2)))


def clip_coef(total_norm: np.float32, max_norm: float) -> np.float32:
    return f32(min(f32(f32(max_norm) / f32(total_norm + f32(1e-6))), f32(1.0)))


def adamw_step(params: List[np.ndarray], grads: List[np.ndarray], state: List[Dict], group_of: Sequence[int],
               groups: Sequence[Dict], max_norm: float = 0.0) -> np.float32:
    """In-place update of params / state[i]['exp_avg', 'exp_avg_sq', 'step'].  groups[j]: lr, betas, eps, weight_decay.
    Returns the total gradient norm (0 when max_norm <= 0)."""
    total = f32(0.0)
    coef = f32(1.0)
    if max_norm and max_norm > 0:
        total = grad_norm(grads)
        coef = clip_coef(total, max_norm)
    for p, g, st, gi in zip(params, grads, state, group_of):
        h = groups[gi]
        b1, b2 = h["betas"]
        st["step"] += 1
        step = st["step"]
        g = (g * coef).astype(f32) if max_norm and max_norm > 0 else g
        p *= f32(1.0 - h["lr"] * h["weight_decay"])
        m, v = st["exp_avg"], st["exp_avg_sq"]
        m += f32(1.0 - b1) * (g - m)
        v *= f32(b2)
        v += (f32(1.0 - b2) * g) * g
        denom = np.sqrt(v) / f32(math.sqrt(1.0 - b2 ** step)) + f32(h["eps"])
        p += f32(-(h["lr"] / (1.0 - b1 ** step))) * (m / denom)
    return total
