"""End-to-end parity of the CUDA forward (through build_model / the C ABI) against the oracle on
the same seeded weights and clips, for both precision tiers.  GPU only.

Tolerances (stated per north_star):
  fp32 tier  : every output within 1e-4 relative (max-norm relative to the tensor's max |value|),
               intermediates within 1e-4 relative L2; decoded event lists identical.
  bf16 tier  : bf16 operands, fp32 accumulation through ~60 layers: relative L2 of every output
               <= 3e-2 and pred_boxes within 2e-2 absolute.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import decode_oracle, sedt_oracle
from sound_event_detection_transformer_b200 import spec, synth
from sound_event_detection_transformer_b200.sedt import build_model

pytestmark = pytest.mark.gpu
torch.set_num_threads(max(1, os.cpu_count() or 1))


def rel_l2(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-12)).item()


def rel_max(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def _cases():
    c1, c2 = spec.config_args("c1"), spec.config_args("c2")
    plain = spec.config_args("c1"); plain.dec_at = False; plain.aux_loss = False
    post = spec.config_args("c1"); post.pre_norm = False
    rag = [synth.synth_clips(1, 500, 64, seed=3)[0], synth.synth_clips(1, 333, 64, seed=4)[0],
           synth.synth_clips(1, 420, 64, seed=5)[0]]
    return {
        "c1_b2": (c1, synth.synth_clips(2, 500, 64, seed=1), 11),
        "c2_b2": (c2, synth.synth_clips(2, 496, 64, seed=2), 12),
        "c1_ragged": (c1, rag, 11),
        "c1_b1": (c1, synth.synth_clips(1, 500, 64, seed=6), 11),
        "c1_plain": (plain, synth.synth_clips(2, 256, 64, seed=7), 13),
        "c1_postnorm": (post, synth.synth_clips(2, 256, 64, seed=8), 14),
    }


def _model(args, seed, precision, use_tc=True):
    args.precision = precision
    args.use_tensor_cores = use_tc
    model, criterion, post = build_model(args)
    model.load_state_dict(synth.synth_state_dict(args, seed), strict=True)
    return model.cuda().eval(), post


def _clips_to_cuda(clips):
    return [c.cuda() for c in clips] if isinstance(clips, list) else clips.cuda()


def _compare(out, ref, tol_fn, tol):
    for k in ("pred_logits", "pred_boxes", "at"):
        if k in ref:
            assert out[k].shape == ref[k].shape, (k, out[k].shape, ref[k].shape)
            assert tol_fn(out[k], ref[k]) < tol, k
    assert len(out.get("aux_outputs", [])) == len(ref.get("aux_outputs", []))
    for a, b in zip(out.get("aux_outputs", []), ref.get("aux_outputs", [])):
        for k in b:
            assert tol_fn(a[k], b[k]) < tol, ("aux", k)


@pytest.mark.parametrize("tag", list(_cases().keys()))
def test_forward_fp32_tier(tag):
    args, clips, seed = _cases()[tag]
    model, post = _model(args, seed, "fp32")
    sd = synth.synth_state_dict(args, seed)
    taps = {}
    ref = sedt_oracle.sedt_forward(sd, args, clips, taps=taps)
    with torch.no_grad():
        out = model(_clips_to_cuda(clips))
    torch.cuda.synchronize()
    _compare(out, ref, rel_max, 1e-4)

    # the oracle itself is pinned to the reference's golden outputs (tests/test_oracle_golden.py);
    # check the CUDA path against those fixtures directly as well
    fx = np.load(os.path.join(GOLDEN, f"sedt_{tag}.npz"))
    for k in ("pred_logits", "pred_boxes", "at"):
        if k in fx:
            assert rel_max(out[k], torch.from_numpy(fx[k])) < 1e-4, k

    # intermediates: layer4 feature map, encoder memory, decoder states
    x, mask = sedt_oracle.nested(clips, None)
    rt = model.runtime()
    padded = bool(mask.any())
    res = rt.forward(x.cuda(), mask.cuda() if padded else None, want_memory=True, want_feat=True)
    torch.cuda.synchronize()
    assert rel_l2(res["feat"].permute(0, 3, 1, 2), taps["layer4"]) < 1e-4
    assert rel_l2(res["memory"], taps["memory"]) < 1e-4
    assert rel_l2(res["hs"], taps["hs"]) < 1e-4

    # decoded events identical (north_star): PostProcess + BoxEncoder.decode_strong restatement
    ev_path = os.path.join(GOLDEN, f"events_{tag}.json")
    if os.path.exists(ev_path):
        gold = json.load(open(ev_path))
        B = out["pred_logits"].shape[0]
        sizes = torch.full((B,), 10.0, device="cuda")
        tags = (out["at"].reshape(B, -1) > 0.5).long()
        names = [f"class{i}" for i in range(10)]
        total = 0
        for at_m in (1, 2, 3):
            res = post["bbox"]({k: v.clone() for k, v in out.items() if k != "aux_outputs"}, sizes, tags, at_m)
            for clip_res, clip_gold in zip(res, gold[str(at_m)]):
                ev = decode_oracle.decode_strong({k: v.cpu().numpy() for k, v in clip_res.items()}, names, 0.5)
                assert [e[0] for e in ev] == [e[0] for e in clip_gold]
                for e, g in zip(ev, clip_gold):
                    assert abs(float(e[1]) - g[1]) < 1e-3 and abs(float(e[2]) - g[2]) < 1e-3 and abs(float(e[3]) - g[3]) < 1e-4
                total += len(ev)
        assert total > 0


@pytest.mark.parametrize("tag", ["c1_b2", "c2_b2", "c1_ragged", "c1_postnorm"])
@pytest.mark.parametrize("use_tc", [False, True])
def test_forward_bf16_tier(tag, use_tc):
    args, clips, seed = _cases()[tag]
    model, _ = _model(args, seed, "bf16", use_tc)
    ref = sedt_oracle.sedt_forward(synth.synth_state_dict(args, seed), args, clips)
    with torch.no_grad():
        out = model(_clips_to_cuda(clips))
    torch.cuda.synchronize()
    _compare(out, ref, rel_l2, 3e-2)
    assert (out["pred_boxes"].cpu() - ref["pred_boxes"]).abs().max() < 2e-2


def test_forward_bf16_tc_matches_bf16_simt():
    """Same bf16 operands, two engines: only the fp32 accumulation order differs."""
    args, clips, seed = _cases()["c2_b2"]
    m_tc, _ = _model(args, seed, "bf16", True)
    with torch.no_grad():
        a = m_tc(clips.cuda())
    m_sm, _ = _model(args, seed, "bf16", False)
    with torch.no_grad():
        b = m_sm(clips.cuda())
    torch.cuda.synchronize()
    for k in ("pred_logits", "pred_boxes", "at"):
        assert rel_l2(a[k], b[k]) < 5e-3, k


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 3e-2)])
def test_spsedt_forward(precision, tol):
    args = spec.config_args("c5")
    model, _ = _model(args, 15, precision)
    sd = synth.synth_state_dict(args, 15)
    x = synth.synth_clips(2, 496, 64, seed=9)
    patches = synth.synth_patches(2, 10, 128, 64, seed=9)
    mask = torch.zeros(2, 496, 64, dtype=torch.bool)
    ref = sedt_oracle.spsedt_forward(sd, args, x, mask, patches)
    with torch.no_grad():
        out = model((x.cuda(), mask.cuda()), patches.cuda())
    torch.cuda.synchronize()
    fn = rel_max if precision == "fp32" else rel_l2
    for k in ("pred_logits", "pred_boxes", "pred_feature", "gt_feature"):
        assert out[k].shape == ref[k].shape, k
        assert fn(out[k], ref[k]) < tol, k
    for a, b in zip(out["aux_outputs"], ref["aux_outputs"]):
        for k in ("pred_logits", "pred_boxes", "pred_feature"):
            assert fn(a[k], b[k]) < tol, ("aux", k)


def test_forward_requires_eval_or_no_grad():
    args = spec.config_args("c1")
    model, _ = _model(args, 11, "fp32")
    model.train()
    with pytest.raises(NotImplementedError):
        model(synth.synth_clips(1, 128, 64).cuda())


def test_weights_repacked_after_update():
    args = spec.config_args("c1")
    model, _ = _model(args, 11, "fp32")
    x = synth.synth_clips(1, 128, 64, seed=1).cuda()
    with torch.no_grad():
        a = model(x)["pred_logits"].clone()
        model.class_embed.bias.add_(1.0)
        b = model(x)["pred_logits"]
    torch.cuda.synchronize()
    assert torch.allclose(b, a + 1.0, atol=1e-5)


def test_cuda_graph_replay_is_bit_identical():
    args, clips, seed = _cases()["c2_b2"]
    model, _ = _model(args, seed, "bf16")
    x = clips.cuda()
    with torch.no_grad():
        ref = {k: v.clone() for k, v in model(x).items() if k != "aux_outputs"}
        model.use_cuda_graph = True
        for _ in range(3):                                   # capture, then replays
            out = model(x)
        torch.cuda.synchronize()
        for k in ref:
            assert torch.equal(out[k], ref[k]), k
        x2 = synth.synth_clips(2, 496, 64, seed=99).cuda()   # new data through the same graph
        out2 = {k: v.clone() for k, v in model(x2).items() if k != "aux_outputs"}
        model.use_cuda_graph = False
        ref2 = model(x2)
        torch.cuda.synchronize()
        for k in out2:
            assert torch.equal(out2[k], ref2[k]), k
