"""SetCriterion (sedt/sedt.py:134-352) on the GPU matcher against golden losses / gradients produced by the
reference's own SetCriterion (tests/golden/make_golden.py: run_criterion).  GPU only (the matcher has no CPU path)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from sound_event_detection_transformer_b200 import spec, synth
from sound_event_detection_transformer_b200.sedt import build_model

pytestmark = pytest.mark.gpu


def _run(tag, cfg, path):
    fx = np.load(os.path.join(GOLDEN, f"criterion_{tag}.npz"))
    B, kmin, kmax, seed = [int(v) for v in fx["meta"]]
    fine_tune, normalize, fl, rng_seed = [int(v) for v in fx["flags"]] if "flags" in fx else (0, 0, 0, 0)
    args = spec.config_args(cfg)
    _, criterion, _ = build_model(args)
    criterion = criterion.cuda()
    criterion.fused = path == "fused"                 # one sedt_set_criterion launch (default)
    if path == "per_clip":                            # neither fused nor batched: the reference's per-clip structure
        criterion._batched_ok = lambda *a, **k: False
    outputs, targets = synth.synth_criterion_case(B, args.num_queries, args.num_classes, args.dec_layers, kmin, kmax, seed)

    def dev(t):
        return t.cuda().requires_grad_(True)
    outputs = {"pred_logits": dev(outputs["pred_logits"]), "pred_boxes": dev(outputs["pred_boxes"]), "at": dev(outputs["at"]),
               "aux_outputs": [{k: dev(v) for k, v in a.items()} for a in outputs["aux_outputs"]]}
    leaves = [outputs["pred_logits"], outputs["pred_boxes"], outputs["at"]]
    for a in outputs["aux_outputs"]:
        leaves += [a["pred_logits"], a["pred_boxes"]]
    for t in targets:
        t["labels"], t["boxes"] = t["labels"].cuda(), t["boxes"].cuda()
    torch.manual_seed(rng_seed)
    losses, indices = criterion(outputs, np.array(targets, dtype=object), None, slice(B), fine_tune=bool(fine_tune),
                                normalize=bool(normalize), fl=bool(fl))
    wd = criterion.weight_dict
    total = sum(losses[k] * wd[k] for k in losses if k in wd)
    total.backward()
    names = [str(n) for n in fx["loss_names"]]
    assert sorted(losses) == names
    for n, v in zip(names, fx["loss_values"]):
        assert abs(float(losses[n]) - v) <= 2e-5 * max(1.0, abs(v)), (n, float(losses[n]), v)
    assert abs(float(total) - float(fx["total"])) <= 2e-5 * abs(float(fx["total"]))
    for i, t in enumerate(leaves):
        ref = torch.from_numpy(fx[f"grad_{i}"])
        assert (t.grad.cpu() - ref).abs().max() <= 1e-6 + 1e-4 * ref.abs().max(), i
    assert len(indices) == B
    if not fine_tune:
        assert all(len(r) == len(c) == min(args.num_queries, len(t["boxes"])) for (r, c), t in zip(indices, targets))


@pytest.mark.parametrize("path", ["fused", "batched", "per_clip"])
@pytest.mark.parametrize("tag,cfg", [("c2", "c2"), ("c1_edges", "c1")])
def test_set_criterion_matches_reference(tag, cfg, path):
    _run(tag, cfg, path)


@pytest.mark.parametrize("tag", ["v_fl", "v_finetune"])
def test_set_criterion_variants_match_reference(tag):
    """focal losses (fl) and the fine_tune + normalize recipe (train_sedt.py:298-310) against the reference's SetCriterion."""
    _run(tag, "c1", "fused")          # these flags route around the fused kernel by themselves


def test_fused_criterion_strong_and_weak_subsets():
    """strong_mask = first 5 clips, weak labels on the first 8 (weakly labelled clips carry labels but no boxes), 12 clips in
    the batch: the fused kernel against the per-clip torch path (which the golden test above ties to the reference)."""
    args = spec.config_args("c1")
    _, criterion, _ = build_model(args)
    criterion = criterion.cuda()
    B = 12
    outputs, targets = synth.synth_criterion_case(B, args.num_queries, args.num_classes, args.dec_layers, 0, 6, 11)
    for t in targets[5:8]:
        t["boxes"] = t["boxes"][:0]
    results = []
    for path in ("fused", "per_clip"):
        criterion.fused = path == "fused"
        criterion._batched_ok = lambda *a, **k: False
        outs = {"pred_logits": outputs["pred_logits"].cuda().requires_grad_(True),
                "pred_boxes": outputs["pred_boxes"].cuda().requires_grad_(True), "at": outputs["at"].cuda().requires_grad_(True),
                "aux_outputs": [{k: v.cuda().requires_grad_(True) for k, v in a.items()} for a in outputs["aux_outputs"]]}
        tg = np.array([{k: v.cuda() for k, v in t.items()} for t in targets], dtype=object)
        losses, indices = criterion(outs, tg, slice(5, 8), slice(5))
        wd = criterion.weight_dict
        sum(losses[k] * wd[k] for k in losses if k in wd).backward()
        leaves = [outs["pred_logits"], outs["pred_boxes"], outs["at"]] + [v for a in outs["aux_outputs"] for v in a.values()]
        results.append((losses, indices, [t.grad.clone() for t in leaves]))
    (lf, idf, gf), (lp, idp, gp) = results
    assert sorted(lf) == sorted(lp)
    for k in lf:
        assert abs(float(lf[k]) - float(lp[k])) <= 2e-5 * max(1.0, abs(float(lp[k]))), k
    for (r1, c1), (r2, c2) in zip(idf, idp):
        assert torch.equal(r1.cpu(), r2.cpu()) and torch.equal(c1.cpu(), c2.cpu())
    for a, b in zip(gf, gp):
        assert (a - b).abs().max() <= 1e-6 + 1e-4 * b.abs().max()
