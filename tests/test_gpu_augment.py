"""Training-time input transforms on the device (csrc/augment.cu via sound_event_detection_transformer_b200.augment) against
golden outputs of the reference's own TimeMask / FreqMask / FreqShift / Query / mixup_data (tests/golden/make_golden.py) and
against the oracle at the full config-5 size.  Everything here is integer / single-rounding fp32 work: the bar is bit-exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import augment_oracle as ao
from sound_event_detection_transformer_b200 import augment, synth

pytestmark = pytest.mark.gpu


def test_augment_clips_matches_reference_transforms():
    fx = np.load(os.path.join(GOLDEN, "augment_b12.npz"))
    B, T, seed = [int(v) for v in fx["meta"]]
    clips = synth.synth_db_clips([T] * B, 64, seed)
    np.random.seed(4200 + seed)
    params = [augment.draw_augment_params(T, 64, np.random, time_mask_args=(0.0, 0.1, 0.7), freq_mask_args=(0.03, 0.4, 0.7),
                                          freq_shift_args=(0.7, 4, 0.0, 2.0)) for _ in range(B)]
    x = torch.from_numpy(np.stack(clips)).cuda()
    got = augment.augment_clips(x, params).cpu().numpy()
    assert np.array_equal(got, fx["out"])
    assert sum(p["tm_t"] > 0 for p in params) and sum(p["fm_mode"] == 2 for p in params) and sum(p["fs_shift"] != 0 for p in params)


def test_augment_clips_full_size_and_constant_fill():
    B, T = 64, 496
    clips = np.stack(synth.synth_db_clips([T] * B, 64, 77))
    rng = np.random.RandomState(5)
    params, want = [], []
    for b in range(B):
        mode = "mean" if b % 2 == 0 else "constant"
        p = augment.draw_augment_params(T, 64, rng, time_mask_args=(0.0, 0.1, 0.8), freq_mask_args=(0.03, 0.4, 0.8),
                                        freq_shift_args=(0.8, 4, 0.0, 2.0), fill_mode=mode, fill_constant=-3.5)
        params.append(p)
        d = clips[b].copy()
        if p["tm_t"]:
            d[p["tm_t0"]:p["tm_t0"] + p["tm_t"], :] *= np.zeros((p["tm_t"], 64))
        if p["fm_mode"]:
            d[:, p["fm_f0"]:p["fm_f0"] + p["fm_f"]] = np.mean(d[:, p["fm_f0"]:p["fm_f0"] + p["fm_f"]]) if p["fm_mode"] == 2 else -3.5
        if p["fs_shift"]:
            d = ao.freq_shift(d, p["fs_shift"])
        want.append(d)
    got = augment.augment_clips(torch.from_numpy(clips.copy()).cuda().unsqueeze(1), params).cpu().numpy()[:, 0]
    assert np.array_equal(got, np.stack(want))


@pytest.mark.parametrize("tag", ["p6", "fixed"])
def test_query_patches_match_reference(tag):
    fx = np.load(os.path.join(GOLDEN, f"query_{tag}.npz"))
    B, P, T, seed, fixed = [int(v) for v in fx["meta"]]
    x = synth.synth_clips(B, T, 64, seed=seed)
    boxes = synth.synth_patch_boxes(B, P, seed, fixed_len=(128 / T) if fixed else None)
    got = augment.query_patches(x.cuda(), boxes, bool(fixed)).cpu().numpy()
    assert got.shape == fx["out"].shape
    assert np.array_equal(got, fx["out"])


def test_query_patches_config5_size_matches_oracle():
    """Config 5's pretraining batch shape: 10 patches per clip of 496 frames (a 32-clip sample through the oracle)."""
    B, P, T = 32, 10, 496
    x = synth.synth_clips(B, T, 64, seed=91)
    boxes = synth.synth_patch_boxes(B, P, 91)
    got = augment.query_patches(x.cuda(), boxes).cpu().numpy()
    for b in range(B):
        assert np.array_equal(got[b], ao.query_patches(x[b].numpy(), boxes[b].numpy())), b
    with pytest.raises(ValueError):
        augment.query_patches(x.cuda(), torch.tensor([[[0.0, 0.2]]]).repeat(B, 1, 1))      # leaves the clip: rejected, not sliced empty


class _NT:
    def __init__(self, t):
        self.tensors = t


@pytest.mark.parametrize("tag", ["ss", "strong_only", "weak_mix"])
def test_mixup_data_matches_reference(tag):
    fx = np.load(os.path.join(GOLDEN, f"mixup_{tag}.npz"))
    n_strong, n_weak, n_unl, T, seed, with_weak = [int(v) for v in fx["meta"]]
    x, y = synth.synth_mixup_case(n_strong, n_weak, n_unl, T, 64, seed)
    y = [{k: v.cuda() for k, v in t.items()} for t in y]
    np.random.seed(4300 + seed)
    nt, labels, s_sl, w_sl = augment.mixup_data(_NT(x.cuda()), np.array(y, dtype=object), slice(n_strong),
                                                slice(n_strong, n_strong + n_weak) if with_weak else None,
                                                mix_up_ratio=0.5, max_events=20, alpha=3)
    assert [s_sl.start or 0, s_sl.stop, w_sl.start, w_sl.stop] == fx["slices"].tolist()
    assert np.array_equal(nt.tensors.cpu().numpy(), fx["out"])
    lo = bo = 0
    assert len(labels) == len(fx["n_labels"])
    for lab, nl, nb in zip(labels, fx["n_labels"], fx["n_boxes"]):
        assert np.array_equal(lab["labels"].cpu().numpy().reshape(-1), fx["labels"][lo:lo + nl])
        assert np.array_equal(lab["boxes"].cpu().numpy().reshape(-1, 2), fx["boxes"][bo:bo + nb])
        if "ratio" in lab:
            assert np.array_equal(lab["ratio"].cpu().numpy(), fx["ratio"][lo:lo + nl])
        else:
            assert (fx["ratio"][lo:lo + nl] < 0).all()
        lo += nl; bo += nb


def test_mixup_label_unlabel_rows():
    x1, y1 = synth.synth_mixup_case(8, 0, 0, 32, 64, 61)
    x2, y2 = synth.synth_mixup_case(8, 0, 0, 32, 64, 62)
    np.random.seed(9)
    lam = np.random.beta(3, 3)
    np.random.seed(9)
    y1c, y2c = [{k: v.cuda() for k, v in t.items()} for t in y1], [{k: v.cuda() for k, v in t.items()} for t in y2]
    out, labels = augment.mixup_label_unlabel(_NT(x1.cuda()), _NT(x2.cuda()), y1c, y2c, mix_up_ratio=0.5, max_events=20, alpha=3)
    got = out.tensors.cpu()
    assert got.shape == x2.shape and len(labels) == 8
    for i in range(8):
        took_l1 = i < 4 and "ratio" not in labels[i] and labels[i]["labels"].data_ptr() == y1c[i]["labels"].data_ptr()
        if i < 4 and "ratio" in labels[i]:
            want = np.float32(lam) * x1[i].numpy() + np.float32(1 - lam) * x2[i].numpy()
        elif took_l1:
            want = x1[i].numpy()
        else:
            want = x2[i].numpy()
        assert np.array_equal(got[i].numpy(), want), i
