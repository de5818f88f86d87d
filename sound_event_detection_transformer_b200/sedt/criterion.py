"""SetCriterion and PostProcess with the reference's call contracts
(sedt/sedt.py:134-352 and :355-396).  Matching runs in the CUDA matcher; the
scalar losses themselves are small batched torch expressions over the matched
pairs -- they are the consumer right after the hot
path, not part of it, and get fused kernels together with the training
backward."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

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


class SetCriterion(nn.Module):
    def __init__(self, num_classes, matcher, weight_dict, eos_coef, losses):
        super().__init__()
        self.num_classes, self.matcher, self.weight_dict = num_classes, matcher, weight_dict
        self.eos_coef, self.losses = eos_coef, losses
        empty_weight = torch.ones(num_classes + 1)
        empty_weight[-1] = eos_coef
        self.register_buffer("empty_weight", empty_weight)

    # -- helpers ------------------------------------------------------------
    @staticmethod
    def _pairs(indices, device):
        batch = torch.cat([torch.full_like(src, i) for i, (src, _) in enumerate(indices)]).to(device)
        src = torch.cat([src for src, _ in indices]).to(device)
        return batch, src

    def _layer_losses(self, out, targets, indices, coef, num_boxes, strong_mask, log: bool):
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
            ce = F.cross_entropy(logits.transpose(1, 2), cls, self.empty_weight, reduction="none")
            res["loss_ce"] = (ce * wq).sum() / num_boxes
            if log:
                if matched.numel() == 0:
                    res["class_error"] = torch.zeros([], device=dev)
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

    def _weak_loss(self, outputs, targets, strong_mask, weak_mask):
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
        return {"loss_weak": F.binary_cross_entropy(pred, gt.clamp(0, 1))}

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
                    res["class_error"] = torch.zeros([], device=dev)
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
        if fl:
            raise NotImplementedError("focal-loss branch (semi-supervised only) is out of scope (SURVEY.md section 2, #9)")
        losses = {}
        indices = None
        if strong_mask is not None and self._batched_ok(targets[strong_mask], fine_tune, normalize):
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
            losses.update(self._layer_losses(outputs, targets, indices, coef, num_boxes, strong_mask, log=True))
        if "weak" in self.losses:
            losses.update(self._weak_loss(outputs, targets, strong_mask, weak_mask))
        if "aux_outputs" in outputs and strong_mask is not None:
            for i, aux in enumerate(outputs["aux_outputs"]):
                sub = {k: v[strong_mask] for k, v in aux.items()}
                sub_idx, sub_coef = self.matcher(sub, targets[strong_mask], fl=fl)
                part = self._layer_losses(aux, targets, sub_idx, sub_coef, num_boxes, strong_mask, log=False)
                losses.update({f"{k}_{i}": v for k, v in part.items()})
        return losses, indices


class PostProcess(nn.Module):
    """outputs -> per-clip {'scores','labels','boxes'} in seconds (sedt/sedt.py:359-396)."""

    @torch.no_grad()
    def forward(self, outputs, target_sizes, audio_tags=None, at_m=2, is_semi=False, threshold=0.5):
        logits, boxes = outputs["pred_logits"], outputs["pred_boxes"]
        prob = F.softmax(logits, -1)
        nq = prob.shape[1]
        if audio_tags is not None:
            ev = prob[..., :-1]                                   # view: writes go through to prob
            best = ev.argmax(1, keepdim=True)                     # per (clip, class): the strongest query
            tags = audio_tags.to(prob.device)
            if at_m in (2, 3):
                top = ev.gather(1, best)
                lift = top < threshold
                if at_m == 3:
                    lift = lift & tags.bool().unsqueeze(1)
                ev.scatter_(1, best, torch.where(lift, torch.full_like(top, threshold), top))
            if at_m in (1, 2):
                ev.mul_(tags.unsqueeze(1).expand(-1, nq, -1).to(ev.dtype))
        scores, labels = prob[..., :-1].max(-1)
        if not is_semi:
            c, l = boxes.unbind(-1)
            se = torch.stack([c - l / 2, c + l / 2], dim=-1)
            se = se * target_sizes.to(se.device).unsqueeze(-1)[:, None, :]
        else:
            se = boxes
        return [{"scores": s, "labels": lb, "boxes": b} for s, lb, b in zip(scores, labels, se)]
