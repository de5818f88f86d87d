"""CUDA matcher (through the reference-shaped HungarianMatcher and the raw C ABI) against the
oracle, scipy and the golden indices produced by the reference.  GPU only."""
import os

import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from conftest import GOLDEN
from oracle import matcher_oracle
from sound_event_detection_transformer_b200 import spec, synth
from sound_event_detection_transformer_b200.sedt import build_matcher
from sound_event_detection_transformer_b200.sedt.matcher import lsap_batched

pytestmark = pytest.mark.gpu


def _to_cuda(outputs, targets):
    return ({k: v.cuda() for k, v in outputs.items()}, [{k: v.cuda() for k, v in t.items()} for t in targets])


@pytest.mark.parametrize("tag,normalize", [("c3_small", False), ("c3_edges", False), ("urban_q10", False),
                                           ("c3_normalize", True)])
def test_matcher_matches_reference_golden(tag, normalize):
    fx = np.load(os.path.join(GOLDEN, f"matcher_{tag}.npz"))
    B, Q, C, kmin, kmax, seed = [int(v) for v in fx["meta"]]
    outputs, targets = synth.synth_matcher_case(B, Q, C, kmin, kmax, seed)
    matcher = build_matcher(spec.default_args())
    idx, coef = matcher(*_to_cuda(outputs, targets), normalize=normalize)
    assert all(not r.is_cuda and r.dtype == torch.int64 and c.dtype == torch.int64 for r, c in idx)   # matcher.py:97
    assert np.array_equal(np.asarray([len(r) for r, _ in idx], np.int32), fx["counts"])
    assert np.array_equal(torch.cat([r for r, _ in idx]).numpy(), fx["rows"])
    assert np.array_equal(torch.cat([c for _, c in idx]).numpy(), fx["cols"])
    oidx, ocoef = matcher_oracle.hungarian_matcher({k: v.numpy() for k, v in outputs.items()},
                                                   [{k: v.numpy() for k, v in t.items()} for t in targets],
                                                   normalize=normalize)
    for cf, ocf in zip(coef, ocoef):
        assert cf.dtype == torch.float32 and np.allclose(cf.numpy(), ocf)


@pytest.mark.parametrize("tag", ["v_fl", "v_finetune", "v_finetune_q10"])
def test_matcher_variants_match_reference_golden(tag):
    """fl (focal class cost), fine_tune (relaxation with the host generator seeded like the reference run) and
    normalize, bit-exact against the reference's own outputs (sedt/matcher.py:77-82,99-133)."""
    fx = np.load(os.path.join(GOLDEN, f"matcher_{tag}.npz"))
    B, Q, C, kmin, kmax, seed, fine_tune, normalize, fl, rng_seed = [int(v) for v in fx["meta"]]
    args = spec.default_args()
    args.epsilon, args.alpha = [float(v) for v in fx["fmeta"]]
    outputs, targets = synth.synth_matcher_case(B, Q, C, kmin, kmax, seed)
    matcher = build_matcher(args)
    torch.manual_seed(rng_seed)
    idx, coef = matcher(*_to_cuda(outputs, targets), fine_tune=bool(fine_tune), normalize=bool(normalize), fl=bool(fl))
    assert np.array_equal(np.asarray([len(r) for r, _ in idx], np.int32), fx["counts"])
    assert np.array_equal(torch.cat([r for r, _ in idx]).numpy(), fx["rows"])
    assert np.array_equal(torch.cat([c for _, c in idx]).numpy(), fx["cols"])
    assert np.allclose(torch.cat(coef).numpy(), fx["coef"])


def test_fine_tune_without_targets_raises_like_the_reference():
    outputs, targets = synth.synth_matcher_case(4, 20, 10, 0, 0, seed=1)
    matcher = build_matcher(spec.default_args())
    with pytest.raises(IndexError):
        matcher(*_to_cuda(outputs, targets), fine_tune=True)


def test_matcher_cost_matrix_matches_oracle():
    outputs, targets = synth.synth_matcher_case(97, 20, 10, 0, 12, seed=11)
    matcher = build_matcher(spec.default_args())
    o, t = _to_cuda(outputs, targets)
    rows, cols, n, cost = matcher.match(o["pred_logits"], o["pred_boxes"], t, return_cost=True)
    cost = cost.cpu().numpy()
    prob = matcher_oracle.softmax_f32(outputs["pred_logits"].numpy())
    for b, tg in enumerate(targets):
        k = len(tg["boxes"])
        ref = matcher_oracle.cost_block(prob[b], outputs["pred_boxes"][b].numpy(), tg["labels"].numpy(), tg["boxes"].numpy())
        assert k == 0 or np.abs(cost[b, :, :k] - ref).max() <= 2e-6          # expf vs numpy exp: a few ulp on the class term
        # the solve on the device's own cost block is scipy's solve, bit for bit
        r, c = linear_sum_assignment(cost[b, :, :k])
        assert np.array_equal(rows[b, :n[b]].cpu().numpy(), r) and np.array_equal(cols[b, :n[b]].cpu().numpy(), c)


@pytest.mark.parametrize("Q,K", [(20, 10), (20, 28), (10, 10), (32, 32), (40, 17), (17, 40), (100, 64), (5, 128)])
def test_lsap_bit_exact_vs_scipy_random(Q, K):
    rng = np.random.default_rng(Q * 1000 + K)
    B = 64
    cost = rng.standard_normal((B, Q, K)).astype(np.float32)
    sizes = rng.integers(0, K + 1, size=B)
    sizes[0], sizes[1] = 0, K
    rows, cols, counts, status = lsap_batched(torch.from_numpy(cost).cuda(), sizes.tolist())
    assert status == 0
    rows, cols, counts = rows.cpu().numpy(), cols.cpu().numpy(), counts.cpu().numpy()
    for b in range(B):
        r, c = linear_sum_assignment(cost[b, :, :sizes[b]])
        assert counts[b] == len(r)
        assert np.array_equal(rows[b, :len(r)], r) and np.array_equal(cols[b, :len(r)], c)
        assert (rows[b, len(r):] == -1).all()


@pytest.mark.parametrize("Q,K", [(20, 10), (8, 8), (6, 15), (20, 28), (40, 40)])
def test_lsap_bit_exact_vs_scipy_ties(Q, K):
    """Integer costs with heavy ties: the tie rule replays scipy's remaining[] order."""
    rng = np.random.default_rng(17)
    B = 128
    cost = rng.integers(0, 3, size=(B, Q, K)).astype(np.float32)
    cost[0] = 1.0                                    # constant matrix (scipy #11602)
    rows, cols, counts, status = lsap_batched(torch.from_numpy(cost).cuda(), [K] * B)
    assert status == 0
    rows, cols = rows.cpu().numpy(), cols.cpu().numpy()
    for b in range(B):
        r, c = linear_sum_assignment(cost[b])
        assert np.array_equal(rows[b, :len(r)], r) and np.array_equal(cols[b, :len(r)], c)


def test_matcher_nan_raises_like_scipy():
    outputs, targets = synth.synth_matcher_case(4, 20, 10, 3, 5, seed=2)
    outputs["pred_boxes"][2, 3, 0] = float("nan")
    matcher = build_matcher(spec.default_args())
    with pytest.raises(ValueError):
        matcher(*_to_cuda(outputs, targets))


def test_matcher_empty_batch_targets():
    outputs, _ = synth.synth_matcher_case(5, 20, 10, 0, 0, seed=2)
    targets = [{"labels": torch.zeros(0, dtype=torch.int64), "boxes": torch.zeros(0, 2)} for _ in range(5)]
    matcher = build_matcher(spec.default_args())
    idx, coef = matcher(*_to_cuda(outputs, targets))
    assert all(len(r) == 0 and len(c) == 0 for r, c in idx) and all(len(c) == 0 for c in coef)


def test_matcher_full_size_properties():
    """Config 3 at full size (8192 clips): permutation validity and optimality vs scipy on a sample."""
    B, Q = 8192, 20
    outputs, targets = synth.synth_matcher_case(B, Q, 10, 0, 10, seed=3)
    matcher = build_matcher(spec.default_args())
    o, t = _to_cuda(outputs, targets)
    rows, cols, n, cost = matcher.match(o["pred_logits"], o["pred_boxes"], t, return_cost=True)
    rows, cols, cost = rows.cpu().numpy(), cols.cpu().numpy(), cost.cpu().numpy()
    for b in range(B):
        k = len(targets[b]["boxes"])
        assert n[b] == min(Q, k)
        r, c = rows[b, :n[b]], cols[b, :n[b]]
        assert np.all(np.diff(r) > 0) and len(set(c.tolist())) == len(c) and (c >= 0).all() and (c < max(k, 1)).all()
    for b in range(0, B, 37):
        k = len(targets[b]["boxes"])
        r, c = linear_sum_assignment(cost[b, :, :k])
        assert np.array_equal(rows[b, :n[b]], r) and np.array_equal(cols[b, :n[b]], c)


def test_packed_targets_and_lazy_indices_match_the_list_api():
    """matcher(outputs, pack_targets(...)) == matcher(outputs, list of dicts): same pairs, per-clip items are views of the two
    [B, Q] result matrices created on demand (sedt/matcher.py:92-97 contract: (rows ascending, cols), CPU int64)."""
    from sound_event_detection_transformer_b200.sedt.matcher import MatchIndices
    outputs, targets = synth.synth_matcher_case(300, 20, 10, 0, 24, seed=9)        # incl. K = 0 and K > Q
    o = {k: v.cuda() for k, v in outputs.items()}
    t = [{k: v.cuda() for k, v in tg.items()} for tg in targets]
    matcher = build_matcher(spec.default_args())
    idx_list, coef_list = matcher(o, t)
    packed = matcher.pack_targets(t, o["pred_logits"].device)
    idx, coef = matcher(o, packed)
    assert isinstance(idx, MatchIndices) and len(idx) == 300 and len(coef) == 300
    for i in range(300):
        (r, c), (r2, c2) = idx[i], idx_list[i]
        assert r.dtype == torch.int64 and not r.is_cuda and torch.equal(r, r2) and torch.equal(c, c2)
        assert len(r) == min(20, len(targets[i]["boxes"])) and torch.equal(coef[i], torch.ones(len(r)))
    assert [len(r) for r, _ in idx[5:9]] == [len(idx[i][0]) for i in range(5, 9)] and len(idx[-1][0]) == len(idx[299][0])
    assert isinstance(coef, list) and float(torch.cat(coef).sum()) == sum(len(r) for r, _ in idx_list)
    matcher.device_indices = True
    idx_dev, _ = matcher(o, packed)
    assert idx_dev[3][0].is_cuda and torch.equal(idx_dev[3][0].cpu(), idx[3][0])
