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


def _run(tag, cfg, batched):
    fx = np.load(os.path.join(GOLDEN, f"criterion_{tag}.npz"))
    B, kmin, kmax, seed = [int(v) for v in fx["meta"]]
    args = spec.config_args(cfg)
    _, criterion, _ = build_model(args)
    criterion = criterion.cuda()
    if not batched:
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
    losses, indices = criterion(outputs, np.array(targets, dtype=object), None, slice(B))
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
    assert len(indices) == B and all(len(r) == len(c) == min(args.num_queries, len(t["boxes"]))
                                     for (r, c), t in zip(indices, targets))


@pytest.mark.parametrize("batched", [True, False])
@pytest.mark.parametrize("tag,cfg", [("c2", "c2"), ("c1_edges", "c1")])
def test_set_criterion_matches_reference(tag, cfg, batched):
    _run(tag, cfg, batched)
