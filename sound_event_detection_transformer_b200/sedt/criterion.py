"""SetCriterion and PostProcess with the reference's call contracts
(sedt/sedt.py:134-352 and :355-396).  Matching runs in the CUDA matcher; the
scalar losses themselves are small batched torch expressions over the matched
pairs -- they are the consumer right after the hot
path, not part of it, and get fused kernels together with the training
backward."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple  # noqa: F401

import torch
import torch.nn.functional as F
from torch import nn


def _se(boxes: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(center, width) -> (onset, offset); utilities/box_ops.py:9-19."""
    c, l = boxes.unbind(-1)
    return c - l / 2, c + l / 2


def paired_l1_giou(src: torch.Tensor, tgt: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-pair L1 on (s,0,e,1) and 1-D GIoU of matched intervals: the diagonal of the
    reference's N x N generalized_box_iou (sedt/sedt.py:249-254) without forming the matrix."""
    s1, e1 = _se(src)
    s2, e2 = _se(tgt)
    l1 = (s1 - s2).abs() + (e1 - e2).abs()
    inter = (torch.min(e1, e2) - torch.max(s1, s2)).clamp(min=0)
    union = (e1 - s1) + (e2 - s2) - inter
    enc = (torch.max(e1, e2) - torch.min(s1, s2)).clamp(min=0)
    giou = inter / union - (enc - union) / enc
    return l1, giou


_LOSS_COL = {"loss_ce": 0, "loss_bbox": 1, "loss_giou": 2, "class_error": 3, "cardinality_error": 4, "loss_weak": 5}


def _common_base(ts: Sequence[torch.Tensor]) -> Optional[torch.Tensor]:
    """The [L, ...] tensor the per-layer outputs are select(0, l) views of (the native forward writes all decoder
    layers into one buffer), or None."""
    base = getattr(ts[0], "_base", None)
    if base is None or base.dim() != ts[0].dim() + 1 or base.shape[0] != len(ts) or not base.is_contiguous():
        return None
    step = base.stride(0) * base.element_size()
    for i, t in enumerate(ts):
        if getattr(t, "_base", None) is not base or t.shape != base.shape[1:] or t.data_ptr() != base.data_ptr() + i * step \
                or not t.is_contiguous():
            return None
    return base


ALPHA_FL, GAMMA_FL = 0.5, 1.0          # config.py:71-72


def sigmoid_focal_loss(inputs, targets, weight=None, alpha=ALPHA_FL, gamma=GAMMA_FL):
    """Per-query focal loss on the class logits (sedt/sedt.py:412-421; the semi-supervised recipe's `fl` branch)."""
    prob = inputs.sigmoid()
    ce = F.binary_cross_entropy_with_logits(inputs, targets, pos_weight=weight, reduction="none")
    p_t = prob * targets + (1 - prob) * (1 - targets)
    loss = ce * (1 - p_t) ** gamma
    if alpha >= 0:
        loss = (alpha * targets + (1 - alpha) * (1 - targets)) * loss
    return loss.sum(2)


def weak_focal_loss(prob, targets, alpha=ALPHA_FL, gamma=GAMMA_FL):
    """Focal variant of the clip-level tagging loss (sedt/sedt.py:424-433)."""
    ce = F.binary_cross_entropy(prob, targets, reduction="none")
    p_t = prob * targets + (1 - prob) * (1 - targets)
    loss = ce * (1 - p_t) ** gamma
    if alpha >= 0:
        loss = (alpha * targets + (1 - alpha) * (1 - targets)) * loss
    return loss.sum(1).mean()


class _FusedSetLoss(torch.autograd.Function):
    """sedt_set_criterion: matcher + losses + gradients of all decoder layers in two launches.  Outputs: one 0-dim
    tensor per requested (layer, loss); backward scales the stored gradients by the incoming scalars."""

    @staticmethod
    def forward(ctx, logits, boxes, at, crit, pk, wanted):
        from .. import _lib
        lib = _lib.load()
        L, B, Q, C1 = logits.shape
        dev = logits.device
        lg = logits.detach().to(torch.float32).contiguous()
        bx = boxes.detach().to(torch.float32).contiguous()
        Bs, Bw = pk["Bs"], (pk["Bw"] if at is not None else 0)
        at2 = at.detach().to(torch.float32).reshape(-1, C1 - 1).contiguous() if at is not None else None
        rows = torch.empty(L, Bs, Q, dtype=torch.int64, device=dev)
        cols = torch.empty(L, Bs, Q, dtype=torch.int64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        partials = torch.empty(L, B, 8, dtype=torch.float32, device=dev)
        vals = torch.empty(L, 8, dtype=torch.float32, device=dev)
        g_logits, g_l1, g_giou = torch.empty_like(lg), torch.empty_like(bx), torch.empty_like(bx)
        g_at = torch.zeros_like(at2) if at2 is not None else None
        m = crit.matcher
        with torch.cuda.device(dev):
            _lib.check(lib.sedt_set_criterion(
                lg.data_ptr(), bx.data_ptr(), _lib.ptr(at2) or None, pk["labels"].data_ptr(), pk["boxes"].data_ptr(),
                pk["offsets"].data_ptr(), pk["n_tgt"].data_ptr(), pk["wl_labels"].data_ptr(), pk["wl_offsets"].data_ptr(),
                L, B, Bs, Bw, Q, C1, pk["kmax"], float(m.cost_class), float(m.cost_bbox), float(m.cost_giou),
                float(crit.eos_coef), float(pk["total"]), rows.data_ptr(), cols.data_ptr(), status.data_ptr(),
                partials.data_ptr(), vals.data_ptr(), g_logits.data_ptr(), g_l1.data_ptr(), g_giou.data_ptr(),
                _lib.ptr(g_at) or None, _lib.current_stream()))
        ctx.g = (g_logits, g_l1, g_giou, g_at)
        ctx.wanted, ctx.at_shape, ctx.L = wanted, (None if at is None else at.shape), L
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(rows, cols, status)
        outs = tuple(vals[l, j] for l, j in wanted)
        return (rows, cols, status) + outs

    @staticmethod
    def backward(ctx, _r, _c, _s, *gs):
        g_logits, g_l1, g_giou, g_at = ctx.g
        L = ctx.L
        dev = g_logits.device
        zero = None
        coef = [[None] * 6 for _ in range(L)]
        for (l, j), g in zip(ctx.wanted, gs):
            coef[l][j] = g
        if all(g is None for g in gs):
            return None, None, None, None, None, None
        zero = torch.zeros([], dtype=torch.float32, device=dev)
        cm = torch.stack([coef[l][j] if coef[l][j] is not None else zero for l in range(L) for j in (0, 1, 2)]).view(L, 3, 1, 1, 1)
        d_logits = g_logits * cm[:, 0]
        d_boxes = g_l1 * cm[:, 1] + g_giou * cm[:, 2]
        d_at = None
        if g_at is not None and coef[L - 1][5] is not None:
            d_at = g_at * coef[L - 1][5]
            n = 1
            for v in ctx.at_shape:
                n *= v
            if d_at.numel() != n:                     # weak loss over the first Bw clips only
                full = torch.zeros(n // d_at.shape[1], d_at.shape[1], dtype=d_at.dtype, device=dev)
                full[:d_at.shape[0]] = d_at
                d_at = full
            d_at = d_at.view(ctx.at_shape)
        return d_logits, d_boxes, d_at, None, None, None


class SetCriterion(nn.Module):
    def __init__(self, num_classes, matcher, weight_dict, eos_coef, losses):
        super().__init__()
        self.num_classes, self.matcher, self.weight_dict = num_classes, matcher, weight_dict
        self.eos_coef, self.losses = eos_coef, losses
        empty_weight = torch.ones(num_classes + 1)
        empty_weight[-1] = eos_coef
        self.register_buffer("empty_weight", empty_weight)
        # one fused matcher + loss + gradient kernel for the supervised default path (csrc/matcher.cu: set_criterion_kernel);
        # False keeps the batched torch expressions below (same values; used by the parity tests as a cross-check)
        self.fused = True

    # -- fused path: sedt_set_criterion ------------------------------------------------------------------------------
    def _fused_ok(self, outputs, targets, strong_mask, weak_mask, fine_tune, normalize) -> bool:
        if not self.fused or fine_tune or normalize or not hasattr(self.matcher, "pack_targets"):
            return False
        if not set(self.losses) <= {"labels", "boxes", "cardinality", "weak"} or not {"labels", "boxes"} <= set(self.losses):
            return False
        for m in (strong_mask, weak_mask):
            if m is not None and not (isinstance(m, slice) and m.step in (None, 1)):
                return False
        if strong_mask is None or strong_mask.start not in (None, 0) or strong_mask.stop is None:
            return False
        lg = outputs["pred_logits"]
        if not lg.is_cuda or lg.dtype != torch.float32 or lg.dim() != 3:
            return False
        return not any("ratio" in t for t in targets)

    def _pack_fused(self, targets, Bs, Bw, Q, dev) -> dict:
        tg = list(targets)
        pk = self.matcher.pack_targets(tg[:Bs], dev)
        sizes = pk["sizes"]
        pk["n"] = [min(Q, k) for k in sizes]
        pk["total"] = sum(pk["n"])
        pk["kmax"] = max(sizes) if sizes else 0
        pk["Bs"], pk["Bw"] = Bs, Bw
        nl = [int(len(t["labels"])) for t in tg]
        pk["n_tgt"] = torch.tensor(nl, dtype=torch.float32).to(dev, non_blocking=True)
        if Bw <= Bs and nl[:Bw] == sizes[:Bw]:
            pk["wl_labels"], pk["wl_offsets"] = pk["labels"], pk["offsets"]        # same table: every label has a box
        else:
            labs = [torch.as_tensor(t["labels"]).reshape(-1) for t in tg[:Bw] if len(t["labels"])]
            pk["wl_labels"] = (torch.cat(labs).to(dev, torch.int64) if labs else torch.zeros(1, dtype=torch.int64, device=dev))
            off = [0]
            for k in nl[:Bw]:
                off.append(off[-1] + k)
            pk["wl_offsets"] = torch.tensor(off, dtype=torch.int32).to(dev, non_blocking=True)
        return pk

    def _forward_fused(self, outputs, targets, strong_mask, weak_mask):
        aux = outputs.get("aux_outputs", [])
        lts = [a["pred_logits"] for a in aux] + [outputs["pred_logits"]]
        bts = [a["pred_boxes"] for a in aux] + [outputs["pred_boxes"]]
        logits = _common_base(lts)
        boxes = _common_base(bts)
        if logits is None or boxes is None:
            logits, boxes = torch.stack(lts), torch.stack(bts)
        L, B, Q, C1 = logits.shape
        dev = logits.device
        Bs = min(int(strong_mask.stop), B)
        use_weak = "weak" in self.losses and "at" in outputs
        Bw = min(int(weak_mask.stop if weak_mask is not None else strong_mask.stop), B) if use_weak else 0
        pk = self._pack_fused(targets, Bs, Bw, Q, dev)
        wanted = []
        names = []
        for l in range(L):
            sfx = "" if l == L - 1 else f"_{l}"
            keys = []
            if "labels" in self.losses:
                keys += ["loss_ce"] + (["class_error"] if l == L - 1 else [])
            if "cardinality" in self.losses:
                keys.append("cardinality_error")
            if "boxes" in self.losses:
                keys += ["loss_bbox", "loss_giou"]
            if l == L - 1 and use_weak:
                keys.append("loss_weak")
            for k in keys:
                wanted.append((l, _LOSS_COL[k]))
                names.append(k + sfx)
        at = outputs["at"] if use_weak else None
        res = _FusedSetLoss.apply(logits, boxes, at, self, pk, tuple(wanted))
        rows, cols, status = res[:3]
        losses = dict(zip(names, res[3:]))
        # reference key order: top layer first (sedt.py:329-351)
        top = {k: v for k, v in losses.items() if not k[-1].isdigit()}
        top.update({k: v for k, v in losses.items() if k[-1].isdigit()})
        # one read-back for the whole step: matcher status + the top layer's index matrices
        packed = torch.cat([status.to(torch.int64), rows[L - 1].flatten(), cols[L - 1].flatten()]).cpu()
        self.matcher.raise_on_status(int(packed[0]))
        r = packed[1:1 + Bs * Q].view(Bs, Q)
        c = packed[1 + Bs * Q:].view(Bs, Q)
        indices = [(r[i, :k], c[i, :k]) for i, k in enumerate(pk["n"])]
        return top, indices

    # -- helpers ------------------------------------------------------------
    @staticmethod
    def _pairs(indices, device):
        batch = torch.cat([torch.full_like(src, i) for i, (src, _) in enumerate(indices)]).to(device)
        src = torch.cat([src for src, _ in indices]).to(device)
        return batch, src

    def _layer_losses(self, out, targets, indices, coef, num_boxes, strong_mask, log: bool, fl: bool = False):
        res = {}
        dev = out["pred_logits"].device
        tg = targets[strong_mask]
        coef_cat = torch.cat(coef).to(dev)
        bi, si = self._pairs(indices, dev)
        if "labels" in self.losses:
            logits = out["pred_logits"][strong_mask]
            matched = torch.cat([t["labels"].to(dev)[j.to(dev)] for t, (_, j) in zip(tg, indices)])
            cls = torch.full(logits.shape[:2], self.num_classes, dtype=torch.int64, device=dev)
            wq = torch.ones(logits.shape[:2], dtype=torch.float32, device=dev)
            cls[bi, si] = matched
            wq[bi, si] = coef_cat
            if fl:                                       # sedt.py:209-217: one-hot over C+1 columns, no-object included
                onehot = F.one_hot(cls, logits.shape[2]).to(logits.dtype)
                ce = sigmoid_focal_loss(logits, onehot, self.empty_weight)
            else:
                ce = F.cross_entropy(logits.transpose(1, 2), cls, self.empty_weight, reduction="none")
            res["loss_ce"] = (ce * wq).sum() / num_boxes
            if log:
                if matched.numel() == 0:                 # utilities/utils.py:566-567: accuracy() is 0 without targets
                    res["class_error"] = torch.full([], 100.0, device=dev)
                else:
                    acc = (logits[bi, si].argmax(-1) == matched).float().mean() * 100.0
                    res["class_error"] = 100 - acc
        if "cardinality" in self.losses:
            with torch.no_grad():
                pl = out["pred_logits"]
                n_tgt = torch.as_tensor([len(v["labels"]) for v in targets], device=dev, dtype=torch.float32)
                n_pred = (pl.argmax(-1) != pl.shape[-1] - 1).sum(1).float()
                res["cardinality_error"] = F.l1_loss(n_pred, n_tgt)
        if "boxes" in self.losses:
            src = out["pred_boxes"][bi, si]
            tgt = torch.cat([t["boxes"].to(dev)[j.to(dev)] for t, (_, j) in zip(targets, indices)], dim=0)
            l1, giou = paired_l1_giou(src, tgt.reshape(-1, 2))
            res["loss_bbox"] = (l1 * coef_cat).sum() / num_boxes
            res["loss_giou"] = ((1 - giou) * coef_cat).sum() / num_boxes
        if "feature" in self.losses:
            gt = out["gt_feature"]
            nb = len(indices)
            gt = gt.view(nb, gt.shape[0] // nb, -1)
            src = F.normalize(out["pred_feature"][bi, si], dim=1)
            tgt = F.normalize(torch.cat([t[j.to(dev)] for t, (_, j) in zip(gt, indices)], dim=0), dim=1)
            res["loss_feature"] = F.mse_loss(src, tgt, reduction="none").sum() / num_boxes
        return res

    def _weak_loss(self, outputs, targets, strong_mask, weak_mask, fl: bool = False):
        if "at" not in outputs:
            return {}
        labeled = slice(weak_mask.stop) if weak_mask is not None else slice(strong_mask.stop)
        pred = outputs["at"][labeled]
        dev = pred.device
        gt = torch.zeros(pred.shape, device=dev)
        # multi-hot clip labels (sedt/sedt.py:169-174), one index_put for the whole batch
        tg = targets[:pred.shape[0]] if not isinstance(targets, (list, tuple)) else list(targets)[:pred.shape[0]]
        sizes = [int(len(t["labels"])) for t in tg]
        if sum(sizes) > 0:
            cols = torch.cat([torch.as_tensor(t["labels"], dtype=torch.int64).reshape(-1) for t in tg]).to(dev)
            rows = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes), output_size=sum(sizes)).to(dev)
            if any("ratio" in t for t in tg):
                vals = torch.cat([t["ratio"].float().reshape(-1).cpu() if "ratio" in t else torch.ones(k)
                                  for t, k in zip(tg, sizes)]).to(dev)
            else:
                vals = torch.ones(sum(sizes), device=dev)
            gt.index_put_((rows, cols), vals, accumulate=True)
        gt = gt.clamp(0, 1)
        return {"loss_weak": weak_focal_loss(pred, gt) if fl else F.binary_cross_entropy(pred, gt)}

    # -- batched path: no per-clip work, no host round trip between the matcher and the losses --------------
    def _batched_ok(self, tg, fine_tune, normalize) -> bool:
        return (not fine_tune and not normalize and "feature" not in self.losses and hasattr(self.matcher, "pack_targets")
                and not any("ratio" in t for t in tg))

    def _pack(self, tg, Q, dev):
        pk = self.matcher.pack_targets(tg, dev)
        n = [min(Q, k) for k in pk["sizes"]]
        total = sum(n)
        nt = torch.tensor(n, dtype=torch.int64)
        bi = torch.repeat_interleave(torch.arange(len(n)), nt, output_size=total)
        starts = torch.cumsum(nt, 0) - nt
        pos = torch.arange(total) - starts[bi]
        pk.update(n=n, total=total, bi=bi.to(dev, non_blocking=True), pos=pos.to(dev, non_blocking=True),
                  off64=pk["offsets"].to(torch.int64))
        return pk

    def _layer_losses_batched(self, out, full_logits, tg, pk, num_boxes, log: bool):
        """Same losses as _layer_losses (sedt/sedt.py:188-261) from the matcher's device-resident index matrices.
        full_logits: pred_logits of the whole batch (the cardinality metric is not restricted to strong_mask)."""
        logits, boxes = out["pred_logits"], out["pred_boxes"]
        dev = logits.device
        rows, cols, _, status = self.matcher.match(logits, boxes, tg, packed=pk, check=False)
        bi, pos = pk["bi"], pk["pos"]
        si, ci = rows[bi, pos], cols[bi, pos]
        gidx = pk["off64"][bi] + ci
        res = {}
        if "labels" in self.losses:
            matched = pk["labels"][gidx]
            cls = torch.full(logits.shape[:2], self.num_classes, dtype=torch.int64, device=dev)
            cls[bi, si] = matched
            ce = F.cross_entropy(logits.transpose(1, 2), cls, self.empty_weight, reduction="none")
            res["loss_ce"] = ce.sum() / num_boxes
            if log:
                if pk["total"] == 0:
                    res["class_error"] = torch.full([], 100.0, device=dev)
                else:
                    res["class_error"] = 100 - (logits[bi, si].argmax(-1) == matched).float().mean() * 100.0
        if "cardinality" in self.losses:
            with torch.no_grad():
                n_pred = (full_logits.argmax(-1) != full_logits.shape[-1] - 1).sum(1).float()
                res["cardinality_error"] = F.l1_loss(n_pred, pk["n_tgt_all"])
        if "boxes" in self.losses:
            l1, giou = paired_l1_giou(boxes[bi, si], pk["boxes"][gidx])
            res["loss_bbox"] = l1.sum() / num_boxes
            res["loss_giou"] = (1 - giou).sum() / num_boxes
        return res, (rows, cols, status)

    # -- reference entry point ---------------------------------------------------
    def forward(self, outputs, targets, weak_mask=None, strong_mask=None, fine_tune=False, normalize=False, fl=False):
        if not fl and self._fused_ok(outputs, targets, strong_mask, weak_mask, fine_tune, normalize):
            return self._forward_fused(outputs, targets, strong_mask, weak_mask)
        losses = {}
        indices = None
        if not fl and strong_mask is not None and self._batched_ok(targets[strong_mask], fine_tune, normalize):
            tg = targets[strong_mask]
            top = {k: v[strong_mask] for k, v in outputs.items() if k != "aux_outputs"}
            dev = top["pred_logits"].device
            pk = self._pack(tg, top["pred_logits"].shape[1], dev)
            num_boxes = torch.as_tensor([float(pk["total"])], dtype=torch.float, device=dev)
            pk["n_tgt_all"] = torch.tensor([len(v["labels"]) for v in targets], dtype=torch.float32).to(dev, non_blocking=True)
            part, (rows, cols, status) = self._layer_losses_batched(top, outputs["pred_logits"], tg, pk, num_boxes, log=True)
            losses.update(part)
            stats = [status]
            if "weak" in self.losses:
                losses.update(self._weak_loss(outputs, targets, strong_mask, weak_mask))
            for i, aux in enumerate(outputs.get("aux_outputs", [])):
                sub = {k: v[strong_mask] for k, v in aux.items()}
                part, (_, _, st) = self._layer_losses_batched(sub, aux["pred_logits"], tg, pk, num_boxes, log=False)
                losses.update({f"{k}_{i}": v for k, v in part.items()})
                stats.append(st)
            # one read-back for the whole step: matcher status of every layer + the top layer's index matrices
            packed = torch.cat([torch.stack(stats).flatten().to(torch.int64), rows.flatten(), cols.flatten()]).cpu()
            for st in packed[:len(stats)].tolist():
                self.matcher.raise_on_status(int(st))
            B, Q = rows.shape
            r = packed[len(stats):len(stats) + B * Q].view(B, Q)
            c = packed[len(stats) + B * Q:].view(B, Q)
            indices = [(r[i, :k], c[i, :k]) for i, k in enumerate(pk["n"])]
            return losses, indices
        if strong_mask is not None:
            top = {k: v[strong_mask] for k, v in outputs.items() if k != "aux_outputs"}
            indices, coef = self.matcher(top, targets[strong_mask], fine_tune=fine_tune, normalize=normalize, fl=fl)
            num_boxes = torch.cat(coef).sum()
            num_boxes = torch.as_tensor([num_boxes], dtype=torch.float, device=outputs["pred_boxes"].device)
            losses.update(self._layer_losses(outputs, targets, indices, coef, num_boxes, strong_mask, log=True, fl=fl))
        if "weak" in self.losses:
            losses.update(self._weak_loss(outputs, targets, strong_mask, weak_mask, fl=fl))
        if "aux_outputs" in outputs and strong_mask is not None:
            for i, aux in enumerate(outputs["aux_outputs"]):
                sub = {k: v[strong_mask] for k, v in aux.items()}
                sub_idx, sub_coef = self.matcher(sub, targets[strong_mask], fl=fl)
                part = self._layer_losses(aux, targets, sub_idx, sub_coef, num_boxes, strong_mask, log=False, fl=fl)
                losses.update({f"{k}_{i}": v for k, v in part.items()})
        return losses, indices


class PostProcess(nn.Module):
    """outputs -> per-clip {'scores','labels','boxes'} in seconds (sedt/sedt.py:359-396).  On CUDA tensors the whole
    batch is one sedt_decode_events launch (softmax, at_m fusion, arg-max, (c,l)->(s,e)); `decode_events` also runs
    BoxEncoder.decode_strong (utilities/BoxEncoder.py:179-226) in the same launch."""

    @staticmethod
    def _run(outputs, target_sizes, audio_tags, at_m, is_semi, threshold, decode_threshold=None):
        from .. import _lib
        lib = _lib.load()
        logits = outputs["pred_logits"].detach().to(torch.float32).contiguous()
        boxes = outputs["pred_boxes"].detach().to(torch.float32).contiguous()
        dev = logits.device
        B, Q, C1 = logits.shape
        sizes = None if is_semi else torch.as_tensor(target_sizes).to(dev, torch.float32).reshape(-1).contiguous()
        if sizes is not None and sizes.numel() != B:
            raise ValueError(f"target_sizes has {sizes.numel()} entries for a batch of {B}")
        tags = None
        if audio_tags is not None:
            tags = audio_tags.to(dev, torch.float32).reshape(B, C1 - 1).contiguous()
        scores = torch.empty(B, Q, dtype=torch.float32, device=dev)
        labels = torch.empty(B, Q, dtype=torch.int64, device=dev)
        se = torch.empty(B, Q, 2, dtype=torch.float32, device=dev)
        ev = None
        if decode_threshold is not None:
            ev = {"cls": torch.empty(B, Q, dtype=torch.int32, device=dev), "onset": torch.empty(B, Q, device=dev),
                  "offset": torch.empty(B, Q, device=dev), "score": torch.empty(B, Q, device=dev),
                  "count": torch.zeros(B, dtype=torch.int32, device=dev)}
        with torch.cuda.device(dev):
            _lib.check(lib.sedt_decode_events(
                logits.data_ptr(), boxes.data_ptr(), _lib.ptr(sizes) or None, _lib.ptr(tags) or None, B, Q, C1, int(at_m),
                float(threshold), int(bool(is_semi)), float(decode_threshold if decode_threshold is not None else 0.0), 0.2,
                scores.data_ptr(), labels.data_ptr(), se.data_ptr(),
                *( [ev[k].data_ptr() for k in ("cls", "onset", "offset", "score", "count")] if ev else [None] * 5),
                _lib.current_stream()))
        return scores, labels, se, ev

    @torch.no_grad()
    def forward(self, outputs, target_sizes, audio_tags=None, at_m=2, is_semi=False, threshold=0.5):
        if not outputs["pred_logits"].is_cuda:
            raise RuntimeError("PostProcess needs CUDA tensors (there is no CPU path)")
        scores, labels, se, _ = self._run(outputs, target_sizes, audio_tags, at_m, is_semi, threshold)
        return [{"scores": s, "labels": lb, "boxes": b} for s, lb, b in zip(scores, labels, se)]

    @torch.no_grad()
    def decode_events(self, outputs, target_sizes, audio_tags=None, at_m=2, threshold=0.5, decode_threshold=0.5,
                      class_names: Optional[Sequence[str]] = None):
        """PostProcess + BoxEncoder.decode_strong(res, decode_threshold) for the whole batch (engine.py:277-291 without
        the per-clip Python loop).  Returns one list per clip of [label, onset, offset, score] (label = class name if
        class_names is given, else the class index), in the reference's order.  One launch, one device->host copy."""
        if not outputs["pred_logits"].is_cuda:
            raise RuntimeError("decode_events needs CUDA tensors (there is no CPU path)")
        _, _, _, ev = self._run(outputs, target_sizes, audio_tags, at_m, False, threshold, decode_threshold)
        B, Q = ev["cls"].shape
        packed = torch.cat([ev["count"].to(torch.float64).reshape(B, 1), ev["cls"].to(torch.float64),
                            ev["onset"].to(torch.float64), ev["offset"].to(torch.float64),
                            ev["score"].to(torch.float64)], dim=1).cpu().numpy()
        out = []
        for row in packed:
            n = int(row[0])
            cls, on, off, sc = row[1:1 + Q], row[1 + Q:1 + 2 * Q], row[1 + 2 * Q:1 + 3 * Q], row[1 + 3 * Q:]
            out.append([[class_names[int(cls[i])] if class_names is not None else int(cls[i]), float(on[i]), float(off[i]),
                         float(sc[i])] for i in range(n)])
        return out

    @torch.no_grad()
    def pseudo_labels(self, tea_outputs, target_sizes, classwise_threshold, del_overlap: bool = True):
        """engine.get_pseudo_labels (engine.py:300-348) for the whole batch in one launch: returns per clip
        {'labels': int64 [n], 'boxes': fp32 [n, 2] (center, width), 'scores': fp32 [n]} on the device, kept queries in the
        reference's order (descending score after the same-class overlap suppression)."""
        from .. import _lib
        lib = _lib.load()
        logits = tea_outputs["pred_logits"]
        if not logits.is_cuda:
            raise RuntimeError("pseudo_labels needs CUDA tensors (there is no CPU path)")
        logits = logits.detach().to(torch.float32).contiguous()
        boxes = tea_outputs["pred_boxes"].detach().to(torch.float32).contiguous()
        dev = logits.device
        B, Q, C1 = logits.shape
        thr = torch.as_tensor(classwise_threshold).to(dev, torch.float32).reshape(-1).contiguous()
        if thr.numel() != C1 - 1:
            raise ValueError(f"classwise_threshold has {thr.numel()} entries for {C1 - 1} classes")
        tags = None
        if "at" in tea_outputs:
            tags = (tea_outputs["at"].reshape(B, C1 - 1) >= thr).to(torch.float32).contiguous()
        min_width = 0.2 / float(torch.as_tensor(target_sizes).reshape(-1)[0])
        labels = torch.empty(B, Q, dtype=torch.int64, device=dev)
        out_boxes = torch.empty(B, Q, 2, dtype=torch.float32, device=dev)
        scores = torch.empty(B, Q, dtype=torch.float32, device=dev)
        counts = torch.zeros(B, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.sedt_pseudo_labels(logits.data_ptr(), boxes.data_ptr(), _lib.ptr(tags) or None, thr.data_ptr(), B, Q, C1,
                                              float(min_width), int(bool(del_overlap)), labels.data_ptr(), out_boxes.data_ptr(),
                                              scores.data_ptr(), counts.data_ptr(), _lib.current_stream()))
        n = counts.tolist()
        return [{"labels": labels[i, :k], "boxes": out_boxes[i, :k], "scores": scores[i, :k]} for i, k in enumerate(n)]
