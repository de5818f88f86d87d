"""On-device PostProcess + BoxEncoder.decode_strong (csrc/decode.cu; sedt/sedt.py:359-396, utilities/BoxEncoder.py:179-226)
against the oracle restatements (which tests/test_oracle_golden.py pins to event lists produced by the reference) and
against the reference's golden event lists themselves."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import decode_oracle, sedt_oracle
from sound_event_detection_transformer_b200.sedt import PostProcess

pytestmark = pytest.mark.gpu
NAMES = [f"class{i}" for i in range(10)]


def _case(seed, B=64, Q=20, C=10, peaky=True):
    """Logits with a few confident queries per clip and clustered intervals, so that thresholds, the 0.2 s filter and the
    per-class overlap suppression all fire."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, Q, C + 1, generator=g)
    if peaky:
        hot = torch.randint(0, 3, (B, Q), generator=g)             # few classes -> same-class overlaps
        boost = (torch.rand(B, Q, generator=g) > 0.4).float() * 6.0
        logits.scatter_add_(2, hot.unsqueeze(-1), boost.unsqueeze(-1))
    centers = torch.rand(B, Q, generator=g)
    widths = torch.rand(B, Q, generator=g) * 0.3
    widths[:, ::5] *= 0.05                                         # some shorter than 0.2 s
    boxes = torch.stack([centers, widths], -1)
    at = torch.rand(B, C, generator=g)
    return {"pred_logits": logits, "pred_boxes": boxes}, at


@pytest.mark.parametrize("at_m", [None, 1, 2, 3])
def test_postprocess_matches_oracle(at_m):
    out, at = _case(1 if at_m is None else at_m)
    tags = None if at_m is None else (at > 0.5).long()
    sizes = torch.full((out["pred_logits"].shape[0],), 10.0)
    ref = sedt_oracle.post_process({k: v.clone() for k, v in out.items()}, sizes, tags, at_m or 2)
    got = PostProcess()({k: v.cuda() for k, v in out.items()}, sizes.cuda(), None if tags is None else tags.cuda(), at_m or 2)
    for a, b in zip(got, ref):
        assert a["labels"].dtype == torch.int64 and torch.equal(a["labels"].cpu(), b["labels"])
        assert torch.allclose(a["scores"].cpu(), b["scores"], atol=2e-7, rtol=1e-6)
        assert torch.allclose(a["boxes"].cpu(), b["boxes"], atol=1e-6)


@pytest.mark.parametrize("at_m", [1, 2, 3])
@pytest.mark.parametrize("Q,C", [(20, 10), (10, 10), (50, 4)])
def test_decode_events_match_oracle(at_m, Q, C):
    out, at = _case(10 * at_m + Q, B=48, Q=Q, C=C)
    names = [f"class{i}" for i in range(C)]
    tags = (at > 0.3).long()
    sizes = torch.full((48,), 10.0)
    post = PostProcess()
    got = post.decode_events({k: v.cuda() for k, v in out.items()}, sizes.cuda(), tags.cuda(), at_m, class_names=names)
    # oracle decode on the kernel's own PostProcess result: the event lists must then be identical, value for value
    res = post({k: v.cuda() for k, v in out.items()}, sizes.cuda(), tags.cuda(), at_m)
    total = suppressed = 0
    for clip_res, ev in zip(res, got):
        r = {k: v.cpu().numpy() for k, v in clip_res.items()}
        want = decode_oracle.decode_strong(r, names, 0.5)
        assert [e[0] for e in ev] == [w[0] for w in want]
        for e, w in zip(ev, want):
            assert e[1] == float(w[1]) and e[2] == float(w[2]) and e[3] == float(w[3])
        total += len(ev)
        kept = int(((r["scores"] >= 0.5) & (r["boxes"][:, 1] - r["boxes"][:, 0] >= 0.2)).sum())
        suppressed += kept - len(ev)
    assert total > 0 and suppressed > 0, "the case must exercise the overlap suppression"


@pytest.mark.parametrize("tag", ["c1_b2", "c2_b2", "c1_ragged", "c1_b1", "c1_postnorm"])
def test_decode_events_match_reference_golden(tag):
    """Event lists the reference's PostProcess + BoxEncoder.decode_strong produced from the reference model's outputs
    (tests/golden/events_*.json); here decoded on the device from the same stored outputs."""
    ev_path = os.path.join(GOLDEN, f"events_{tag}.json")
    fx = np.load(os.path.join(GOLDEN, f"sedt_{tag}.npz"))
    gold = json.load(open(ev_path))
    out = {"pred_logits": torch.from_numpy(fx["pred_logits"]).cuda(), "pred_boxes": torch.from_numpy(fx["pred_boxes"]).cuda()}
    B = out["pred_logits"].shape[0]
    at = torch.from_numpy(fx["at"]).reshape(B, -1)
    tags = (at > 0.5).long().cuda()
    sizes = torch.full((B,), 10.0, device="cuda")
    total = 0
    for at_m in (1, 2, 3):
        got = PostProcess().decode_events(out, sizes, tags, at_m, class_names=NAMES)
        for ev, clip_gold in zip(got, gold[str(at_m)]):
            assert [e[0] for e in ev] == [g[0] for g in clip_gold]
            for e, g in zip(ev, clip_gold):
                assert abs(e[1] - g[1]) < 1e-5 and abs(e[2] - g[2]) < 1e-5 and abs(e[3] - g[3]) < 1e-5
            total += len(ev)
    assert total > 0


@pytest.mark.parametrize("tag", ["q20", "q10_keepall"])
def test_pseudo_labels_match_reference_golden(tag):
    """PostProcess.pseudo_labels (csrc/decode.cu: pseudo_labels_kernel) against engine.get_pseudo_labels of the reference
    (engine.py:300-348; fixture made by tests/golden/make_golden.py: run_pseudo_labels): labels and boxes identical."""
    from sound_event_detection_transformer_b200 import synth
    fx = np.load(os.path.join(GOLDEN, f"pseudo_{tag}.npz"))
    B, Q, C, seed, del_overlap = [int(v) for v in fx["meta"]]
    logits, boxes, at = synth.synth_teacher_case(B, Q, C, seed)
    out = PostProcess().pseudo_labels({"pred_logits": logits.cuda(), "pred_boxes": boxes.cuda(), "at": at.cuda()},
                                      torch.full((B,), 10.0), torch.from_numpy(fx["thr"]), del_overlap=bool(del_overlap))
    assert np.array_equal(np.asarray([len(t["labels"]) for t in out], np.int32), fx["counts"])
    assert np.array_equal(torch.cat([t["labels"] for t in out]).cpu().numpy(), fx["labels"])
    assert np.array_equal(torch.cat([t["boxes"] for t in out]).cpu().numpy(), fx["boxes"])
    assert fx["counts"].sum() > 0


def test_pseudo_labels_match_oracle_large():
    from sound_event_detection_transformer_b200 import synth
    logits, boxes, at = synth.synth_teacher_case(96, 40, 6, 77)
    thr = np.linspace(0.4, 0.6, 6).astype(np.float32)
    want = decode_oracle.pseudo_labels(logits.numpy(), boxes.numpy(), at.numpy(), thr, 10.0, True)
    got = PostProcess().pseudo_labels({"pred_logits": logits.cuda(), "pred_boxes": boxes.cuda(), "at": at.cuda()},
                                      torch.full((96,), 10.0), torch.from_numpy(thr))
    suppressed = 0
    for (wl, wb), g in zip(want, got):
        assert np.array_equal(g["labels"].cpu().numpy(), wl) and np.array_equal(g["boxes"].cpu().numpy(), wb)
    assert sum(len(wl) for wl, _ in want) > 0
