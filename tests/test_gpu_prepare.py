"""Device input pipeline (csrc/prepare.cu; utilities/BoxTransforms.py:454-490 evaluation recipe) against the reference's own
PadOrTrunc / ToTensor / Normalize output (golden prepare_ragged.npz: bit-exact, the float64 normalisation included) and
against the oracle with the dB conversion (log10f vs numpy log10: 2e-6 relative of the dB value before normalisation)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import prepare_oracle
from sound_event_detection_transformer_b200 import synth
from sound_event_detection_transformer_b200.prepare import prepare_clips

pytestmark = pytest.mark.gpu


def test_pad_normalize_bit_exact_vs_reference():
    fx = np.load(os.path.join(GOLDEN, "prepare_ragged.npz"))
    frames, F, seed = [int(v) for v in fx["meta"]]
    clips = [torch.from_numpy(c) for c in synth.synth_db_clips([int(v) for v in fx["lengths"]], F, seed)]
    out = prepare_clips(clips, frames, torch.from_numpy(fx["mean"]), torch.from_numpy(fx["std"]), apply_log=False)
    assert out.shape == (len(clips), 1, frames, F) and out.dtype == torch.float32
    assert np.array_equal(out.cpu().numpy(), fx["out"])


@pytest.mark.parametrize("frames", [496, 500])
def test_full_pipeline_vs_oracle(frames):
    rng = np.random.default_rng(frames)
    lengths = [300, 496, 500, 700, 1, 499]
    raw = [np.abs(rng.standard_normal((t, 64))).astype(np.float32) * rng.uniform(0.01, 5.0) for t in lengths]
    raw[1][3, 5] = 0.0                                           # hits the amin floor and the top_db clip
    mean = rng.standard_normal(64) * 5.0 - 20.0
    std = rng.uniform(5.0, 15.0, 64)
    got = prepare_clips([torch.from_numpy(r) for r in raw], frames, torch.from_numpy(mean), torch.from_numpy(std)).cpu().numpy()
    want = np.stack([prepare_oracle.prepare_clip(r, frames, mean, std) for r in raw])
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-6 * 100.0 / std.min() + 1e-6          # |dB| <= 100 before the normalisation
    # no scaler: the padded dB clip itself
    got = prepare_clips([torch.from_numpy(r) for r in raw], frames).cpu().numpy()
    want = np.stack([prepare_oracle.prepare_clip(r, frames) for r in raw])
    assert np.abs(got - want).max() <= 2e-4 and np.array_equal(got == 0.0, want == 0.0)


def test_cpu_device_raises():
    with pytest.raises(RuntimeError):
        prepare_clips([torch.zeros(4, 64)], 8, device="cpu")
