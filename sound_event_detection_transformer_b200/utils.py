"""Boundary types shared with the reference's callers (engine.py, train_*.py):
NestedTensor and nested_tensor_from_tensor_list with the reference's semantics
(utilities/utils.py:470-492, :526-560): clips are zero-padded to the batch
maximum and `mask` is True on padding."""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor


class NestedTensor(object):
    def __init__(self, tensors: Tensor, mask: Optional[Tensor]):
        self.tensors = tensors
        self.mask = mask
        # host-side knowledge that nothing is padded (lets the runtime skip the per-clip
        # position table without a device sync); None = unknown
        self.unpadded: Optional[bool] = None

    def to(self, device):
        out = NestedTensor(self.tensors.to(device), None if self.mask is None else self.mask.to(device))
        out.unpadded = self.unpadded
        return out

    def cuda(self, non_blocking: bool = True):
        out = NestedTensor(self.tensors.cuda(non_blocking=non_blocking),
                           None if self.mask is None else self.mask.cuda(non_blocking=non_blocking))
        out.unpadded = self.unpadded
        return out

    def decompose(self):
        return self.tensors, self.mask

    def __getitem__(self, i: slice):
        if isinstance(i, slice):
            out = NestedTensor(self.tensors[i], self.mask[i])
            out.unpadded = self.unpadded
            return out
        raise TypeError("NestedTensor only supports slicing")

    def __repr__(self):
        return str(self.tensors)


def nested_tensor_from_tensor_list(tensor_list) -> NestedTensor:
    """list of [C,T,F] clips (or a [B,C,T,F] tensor) -> zero-padded batch + bool mask."""
    if isinstance(tensor_list, Tensor):
        if tensor_list.ndim != 4:
            raise ValueError("not supported")
        b, _, h, w = tensor_list.shape
        nt = NestedTensor(tensor_list, torch.zeros((b, h, w), dtype=torch.bool, device=tensor_list.device))
        nt.unpadded = True
        return nt
    if tensor_list[0].ndim != 3:
        raise ValueError("not supported")
    shapes = [tuple(t.shape) for t in tensor_list]
    max_size = [max(s[i] for s in shapes) for i in range(3)]
    b = len(tensor_list)
    c, h, w = max_size
    dtype, device = tensor_list[0].dtype, tensor_list[0].device
    same = all(s == shapes[0] for s in shapes)
    if same:
        tensor = torch.stack(list(tensor_list))
        mask = torch.zeros((b, h, w), dtype=torch.bool, device=device)
    else:
        tensor = torch.zeros([b] + max_size, dtype=dtype, device=device)
        mask = torch.ones((b, h, w), dtype=torch.bool, device=device)
        for img, pad_img, m in zip(tensor_list, tensor, mask):
            pad_img[: img.shape[0], : img.shape[1], : img.shape[2]].copy_(img)
            m[: img.shape[1], : img.shape[2]] = False
    nt = NestedTensor(tensor, mask)
    nt.unpadded = same
    return nt
