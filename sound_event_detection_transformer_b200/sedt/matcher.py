"""HungarianMatcher with the reference's call contract (sedt/matcher.py:41-133)
running on the GPU: one sedt_matcher launch builds every clip's cost block and
solves it (one warp per clip), replacing the [B*Q, sum K] cross-batch matrix,
the .cpu() copy and the serial scipy loop.  All of the reference's branches are
covered: the focal class cost (fl), the fine_tune relaxation, and the
normalize / mixup-ratio coefficients."""
from __future__ import annotations

import ctypes as C
from collections import Counter
from collections.abc import Sequence as _Seq
from typing import List, Sequence, Tuple

import torch
from torch import nn

from .. import _lib


class MatchIndices(_Seq):
    """The matcher's `indices` result, `indices[i] == (rows_i int64 ascending, cols_i int64)` (sedt/matcher.py:92-97), backed by
    the two padded [B, Q] matrices the kernel writes: clip i's pairs are the first counts[i] entries of row i, so every item
    is a view that is only created when somebody asks for it (8192 clips -> no 16384 small tensors up front)."""

    def __init__(self, rows: torch.Tensor, cols: torch.Tensor, counts: Sequence[int]):
        self.rows, self.cols, self.counts = rows, cols, counts

    def __len__(self):
        return len(self.counts)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        n = self.counts[i]
        return self.rows[i, :n], self.cols[i, :n]


def ones_coef(counts: Sequence[int], q: int) -> List[torch.Tensor]:
    """Coef of the default recipe (sedt/matcher.py:131): ones(len(indices[i])) per clip.  A real list (consumers call
    torch.cat(coef), sedt/sedt.py:322) whose entries are shared: one tensor per distinct length."""
    ones = torch.ones(max(q, 1), dtype=torch.float32)
    cache = [ones[:n] for n in range(max(q, 1) + 1)]
    return [cache[n] for n in counts]


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1, epsilon=0, alpha=100):
        super().__init__()
        self.cost_class, self.cost_bbox, self.cost_giou = cost_class, cost_bbox, cost_giou
        self.epsilon, self.alpha = epsilon, alpha
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"
        self.device_indices = False      # True: keep index tensors on the GPU (skips the D2H copy + sync)
        self.alpha_fl, self.gamma_fl = 0.5, 1.0      # config.py:71-72 (module constants in the reference)

    @torch.no_grad()
    def forward(self, outputs, targets: Sequence[dict], fine_tune=False, normalize=False, fl=False):
        """Returns (indices, Coef): indices[i] = (int64 rows ascending, int64 cols) with
        len = min(num_queries, K_i); Coef[i] fp32 (ones | 1/multiplicity | targets[i]['ratio'])."""
        # targets: the reference's list of dicts, or a pack_targets() result (labels / boxes / offsets already on the device:
        # a training loop packs once per batch and reuses it for the matcher calls of all decoder layers)
        packed = targets if isinstance(targets, dict) and "offsets" in targets else None
        tlist = [None] * len(packed["sizes"]) if packed is not None else targets
        if fine_tune:
            rows, cols, counts, lmin, largmin = self.match(outputs["pred_logits"], outputs["pred_boxes"], tlist, fl=fl,
                                                           want_lmin=True, packed=packed)
        else:
            rows, cols, counts = self.match(outputs["pred_logits"], outputs["pred_boxes"], tlist, fl=fl, packed=packed)
        if not fine_tune and not normalize and (packed is not None or not any("ratio" in t for t in targets)):
            # default recipe: the padded [B, Q] matrices ARE the result (pairs first, sorted by query): one D2H copy of each
            # (reference contract: CPU tensors) or none (device_indices), per-clip views on demand
            if not self.device_indices:
                rows, cols = rows.cpu(), cols.cpu()
            return MatchIndices(rows, cols, counts), ones_coef(counts, rows.shape[1])
        if packed is not None:
            raise ValueError("fine_tune / normalize / ratio matching needs the per-clip target dicts")
        # compact the padded [B,Q] index matrices once on the device, then one split on the host
        valid = rows >= 0
        flat_r, flat_c = rows[valid], cols[valid]
        if not self.device_indices or fine_tune:
            flat_r, flat_c = flat_r.cpu(), flat_c.cpu()
        rs, cs = flat_r.split(counts), flat_c.split(counts)
        idx: List[Tuple[torch.Tensor, torch.Tensor]] = list(zip(rs, cs))
        if fine_tune:
            idx = self._relax(idx, lmin.cpu(), largmin.cpu(), rows.shape[1], [int(len(v["boxes"])) for v in targets])
            rs, cs = [r for r, _ in idx], [c for _, c in idx]
            counts = [len(r) for r in rs]
        coef: List[torch.Tensor] = []
        if normalize or any("ratio" in t for t in targets):
            for i, n in enumerate(counts):
                if normalize:
                    cur = cs[i].tolist()
                    num = Counter(cur)
                    coef.append(torch.tensor([1 / num[j] for j in cur], dtype=torch.float32))
                elif "ratio" in targets[i]:
                    coef.append(targets[i]["ratio"].cpu())
                else:
                    coef.append(torch.ones(n, dtype=torch.float32))
        else:
            coef = list(torch.ones(sum(counts), dtype=torch.float32).split(counts))
        return idx, coef

    def _relax(self, idx1, lmin, largmin, num_queries, sizes):
        """The fine_tune relaxation (sedt/matcher.py:99-121), statement for statement on the kernel's per-query location-cost
        minimum / arg-min: Hungarian pairs survive only where the query's best location cost is below epsilon; other
        queries below epsilon are attached to their nearest target, each kept with probability alpha * num_gt / num_queries
        (torch.rand on the host generator, drawn per clip in the reference's order, so a seeded run reproduces it)."""
        out = []
        for b, (r, c) in enumerate(idx1):
            if sizes[b] == 0:
                raise IndexError("fine_tune matching needs at least one target per clip "
                                 "(the reference's c[i].min(-1) fails on an empty target list, sedt/matcher.py:106)")
            num_gt = len(c)
            reserved = lmin[b] < self.epsilon
            keep = reserved[r] == True                                   # noqa: E712
            r, c = r[keep], c[keep]
            reserved[r] = False
            reserved_index = torch.where(reserved == True)[0]            # noqa: E712
            random_del_index = torch.where(torch.rand(len(reserved_index)) > (self.alpha * num_gt / num_queries))[0]
            reserved[reserved_index[random_del_index]] = False
            out.append((torch.cat([r, torch.arange(num_queries)[reserved]], dim=-1),
                        torch.cat([c, largmin[b][reserved]], dim=-1)))
        return out

    @staticmethod
    def pack_targets(targets: Sequence[dict], dev) -> dict:
        """Concatenate the per-clip target lists once (labels int64 [sumK], boxes fp32 [sumK,2], offsets int32 [B+1]
        on the device); SetCriterion reuses one pack for the matcher calls of all decoder layers."""
        sizes = [int(len(v["boxes"])) for v in targets]
        if sum(sizes) > 0:
            labs = [v["labels"] if v["labels"].shape[0] == k else v["labels"][:k] for v, k in zip(targets, sizes)]
            tgt_ids = torch.cat(labs).reshape(-1).to(dev, torch.int64).contiguous()
            tgt_box = torch.cat([v["boxes"] for v in targets]).reshape(-1, 2).to(dev, torch.float32).contiguous()
        else:
            tgt_ids = torch.zeros(1, dtype=torch.int64, device=dev)
            tgt_box = torch.zeros(1, 2, dtype=torch.float32, device=dev)
        off = [0]
        for k in sizes:
            off.append(off[-1] + k)
        offsets = torch.tensor(off, dtype=torch.int32).to(dev, non_blocking=True)
        return {"sizes": sizes, "labels": tgt_ids, "boxes": tgt_box, "offsets": offsets}

    @torch.no_grad()
    def match(self, pred_logits: torch.Tensor, pred_boxes: torch.Tensor, targets: Sequence[dict],
              return_cost: bool = False, packed: dict = None, check: bool = True, fl: bool = False, want_lmin: bool = False):
        """Raw batched call: returns rows [B,Q] int64, cols [B,Q] int64 on the device (every clip's pairs first,
        sorted by query, then -1 padding) and the per-clip pair counts as a Python list (known on the host:
        min(Q, K_i)).  packed: a pack_targets() result to reuse; check=False skips the status read-back (no
        sync; the status tensor is returned as a fourth value instead)."""
        lib = _lib.load()
        if not pred_logits.is_cuda:
            raise RuntimeError("HungarianMatcher needs CUDA tensors (there is no CPU path)")
        dev = pred_logits.device
        B, Q, C1 = pred_logits.shape
        assert len(targets) == B
        logits = pred_logits.detach().to(torch.float32).contiguous()
        boxes = pred_boxes.detach().to(torch.float32).contiguous()
        if packed is not None:
            sizes, tgt_ids, tgt_box, offsets = packed["sizes"], packed["labels"], packed["boxes"], packed["offsets"]
            kmax = max(sizes) if sizes else 0
        else:
            sizes = [int(len(v["boxes"])) for v in targets]
            kmax = max(sizes) if sizes else 0
        if packed is not None:
            pass
        elif sum(sizes) > 0:
            # matcher.py:69 slices labels to the number of boxes; skip the per-clip slice when they already agree
            labs = [v["labels"] if v["labels"].shape[0] == k else v["labels"][:k] for v, k in zip(targets, sizes)]
            tgt_ids = torch.cat(labs).reshape(-1).to(dev, torch.int64).contiguous()
            tgt_box = torch.cat([v["boxes"] for v in targets]).reshape(-1, 2).to(dev, torch.float32).contiguous()
        else:
            tgt_ids = torch.zeros(1, dtype=torch.int64, device=dev)
            tgt_box = torch.zeros(1, 2, dtype=torch.float32, device=dev)
        if packed is None:
            off = [0]
            for k in sizes:
                off.append(off[-1] + k)
            offsets = torch.tensor(off, dtype=torch.int32).to(dev, non_blocking=True)
        rows = torch.empty(B, Q, dtype=torch.int64, device=dev)
        cols = torch.empty(B, Q, dtype=torch.int64, device=dev)
        counts = torch.empty(B, dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        cost = torch.full((B, Q, max(kmax, 1)), float("nan"), dtype=torch.float32, device=dev) if return_cost else None
        lmin = torch.empty(B, Q, dtype=torch.float32, device=dev) if want_lmin else None
        largmin = torch.empty(B, Q, dtype=torch.int64, device=dev) if want_lmin else None
        with torch.cuda.device(dev):
            _lib.check(lib.sedt_matcher_ex(logits.data_ptr(), boxes.data_ptr(), tgt_ids.data_ptr(), tgt_box.data_ptr(),
                                           offsets.data_ptr(), B, Q, C1, kmax, float(self.cost_class), float(self.cost_bbox),
                                           float(self.cost_giou), int(bool(fl)), float(self.alpha_fl), float(self.gamma_fl),
                                           _lib.ptr(cost) or None, max(kmax, 1), rows.data_ptr(), cols.data_ptr(),
                                           counts.data_ptr(), status.data_ptr(), _lib.ptr(lmin) or None,
                                           _lib.ptr(largmin) or None, _lib.current_stream()))
        n = [min(Q, k) for k in sizes]
        if not check:
            return rows, cols, n, status
        if not self.device_indices or want_lmin:
            self.raise_on_status(int(status.item()))
        if want_lmin:
            return rows, cols, n, lmin, largmin
        if return_cost:
            return rows, cols, n, cost
        return rows, cols, n

    @staticmethod
    def raise_on_status(st: int):
        if st == -4:
            raise ValueError("matrix contains invalid numeric entries")      # scipy's message (matcher.py:95)
        if st != 0:
            raise ValueError("cost matrix is infeasible")


def lsap_batched(cost: torch.Tensor, sizes: Sequence[int]):
    """scipy.optimize.linear_sum_assignment on every cost[b, :, :sizes[b]] in one launch (test hook)."""
    lib = _lib.load()
    dev = cost.device
    B, Q, ld = cost.shape
    cost = cost.to(torch.float32).contiguous()
    off = [0]
    for k in sizes:
        off.append(off[-1] + int(k))
    offsets = torch.tensor(off, dtype=torch.int32, device=dev)
    rows = torch.empty(B, Q, dtype=torch.int64, device=dev)
    cols = torch.empty(B, Q, dtype=torch.int64, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sedt_lsap(cost.data_ptr(), ld, offsets.data_ptr(), B, Q, max([int(k) for k in sizes] + [0]),
                                 rows.data_ptr(), cols.data_ptr(), counts.data_ptr(), status.data_ptr(),
                                 _lib.current_stream()))
    return rows, cols, counts, int(status.item())


def build_matcher(args):
    return HungarianMatcher(cost_class=args.set_cost_class, cost_bbox=args.set_cost_bbox,
                            cost_giou=args.set_cost_giou, epsilon=args.epsilon, alpha=args.alpha)
