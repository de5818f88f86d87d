"""Parity of the configurations the benchmark numbers are quoted on, at their FULL batch sizes and with the DEFAULT
dispatch (launch_conv_tc picks different kernels at B = 256 than at the B <= 3 of tests/test_gpu_forward.py; the fused
blocks only engage where whole rounds of tiles fill the SMs).  Every test asserts which kernels produced the output it
checked (per-kind launch counters of the library).  GPU only.

  config 2 (BASELINE.json configs[1], the headline): SEDT E=6, Q=20, B=256 clips [1,496,64], bf16 tier
      vs the fp32 oracle on all 256 clips (rel-L2 <= 3e-2) and vs the bf16-rounding oracle on a subsample (tight bar:
      only the fp32 accumulation order differs), graph replay bit-identical to the eager launches, decoded events of the
      bf16 tier vs the oracle's (identical on every clip whose decoding is stable under a perturbation of the size of
      the bf16 error; agreement rate over all clips reported and bounded).
  config 5: SP-SEDT forward, B=200 clips + 2000 patches.
  config 4: one E=6, T=496, B=64 training step vs autograd through the bf16-rounding oracle (rel-L2, cosine, projection).
"""
import os

import numpy as np
import pytest
import torch

from oracle import bf16_oracle, decode_oracle, sedt_oracle
from sound_event_detection_transformer_b200 import _lib, spec, synth
from sound_event_detection_transformer_b200.sedt import build_model

pytestmark = pytest.mark.gpu
torch.set_num_threads(max(1, os.cpu_count() or 1))


def rel_l2(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-12)).item()


def _kinds_delta(before):
    now = _lib.kernel_kind_counts()
    return {k: now[k] - before[k] for k in now}


def _flat_outputs(out):
    res = {k: out[k] for k in ("pred_logits", "pred_boxes", "at", "pred_feature", "gt_feature") if k in out}
    for i, aux in enumerate(out.get("aux_outputs", [])):
        for k in ("pred_logits", "pred_boxes", "pred_feature"):
            if k in aux:
                res[f"aux{i}.{k}"] = aux[k]
    return res


@pytest.fixture(scope="module")
def c2_b256():
    """One eager B = 256 forward of the headline configuration + the oracle on the same clips."""
    args = spec.config_args("c2")
    args.precision = "bf16"
    sd = synth.synth_state_dict(args, 12)
    model, _, post = build_model(args)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    clips = synth.synth_clips(256, 496, 64, seed=100)
    x = clips.cuda()
    with torch.no_grad():
        model(x[:2])                               # warm the one-time setup outside the counted pass
        k0 = _lib.kernel_kind_counts()
        out = model(x)
        torch.cuda.synchronize()
    kinds = _kinds_delta(k0)
    out = {k: (v.clone() if torch.is_tensor(v) else [{a: b.clone() for a, b in d.items()} for d in v]) for k, v in out.items()}
    ref = {}
    for i in range(0, 256, 64):                    # chunked: bounds the oracle's host memory
        r = sedt_oracle.sedt_forward(sd, args, clips[i:i + 64])
        for k, v in _flat_outputs(r).items():
            ref.setdefault(k, []).append(v)
    ref = {k: torch.cat(v) for k, v in ref.items()}
    return dict(args=args, sd=sd, model=model, post=post, clips=clips, x=x, out=out, kinds=kinds, ref=ref)


def test_c2_b256_uses_the_benchmarked_kernels(c2_b256):
    k = c2_b256["kinds"]
    print("kernel kinds of one B=256 forward:", k)
    assert k["stem_tc"] == 1
    assert k["conv_tc3_2sm"] > 0, k               # cta_group::2 kernel: K >= 512 layers
    assert k["conv_tc4_ws"] + k["bottleneck_fused"] > 0, k      # short-K weight-stationary kernel / fused bottleneck blocks
    assert k["ffn_fused"] == c2_b256["args"].enc_layers, k
    assert k["attention_tc"] + k["enc_attn_fused"] + k["dec_layer_fused"] > 0, k
    assert k["conv_tc_v1"] == 0, k


def test_c2_b256_matches_fp32_oracle(c2_b256):
    got = _flat_outputs(c2_b256["out"])
    for k, r in c2_b256["ref"].items():
        e = rel_l2(got[k], r)
        print(f"c2 B=256 bf16 vs fp32 oracle {k}: rel-L2 {e:.2e}")
        assert e < 3e-2, (k, e)
    assert (got["pred_boxes"].cpu() - c2_b256["ref"]["pred_boxes"]).abs().max() < 2e-2


def test_c2_b256_matches_bf16_rounding_oracle(c2_b256):
    """Same rounding points on both sides: what is left is fp32 summation order (and the bf16 ulp flips it causes)."""
    idx = list(range(0, 256, 8))                   # 32 clips spread over the batch (every CTA wave is sampled)
    sub = c2_b256["clips"][idx]
    ref = _flat_outputs(bf16_oracle.sedt_forward_bf16(c2_b256["sd"], c2_b256["args"], sub))
    got = _flat_outputs(c2_b256["out"])
    for k, r in ref.items():
        e = rel_l2(got[k][idx], r)
        print(f"c2 B=256 bf16 vs bf16-rounding oracle {k}: rel-L2 {e:.2e}")
        assert e < 4e-3, (k, e)


def test_c2_b256_rows_match_small_batch_and_graph_replay(c2_b256):
    model, x, out = c2_b256["model"], c2_b256["x"], c2_b256["out"]
    with torch.no_grad():
        small = model(x[40:42])
        torch.cuda.synchronize()
        for k in ("pred_logits", "pred_boxes", "at"):       # other kernels at B = 2 (pinned by test_gpu_forward.py): same math
            assert rel_l2(out[k][40:42], small[k]) < 5e-3, k
        model.use_cuda_graph = True
        try:
            for _ in range(2):
                g = model(x)
            torch.cuda.synchronize()
            for k in ("pred_logits", "pred_boxes", "at"):
                assert torch.equal(g[k], out[k]), k
        finally:
            model.use_cuda_graph = False


def _decode(post, outputs, names):
    B = outputs["pred_logits"].shape[0]
    sizes = torch.full((B,), 10.0, device="cuda")
    tags = (outputs["at"].reshape(B, -1) > 0.5).long()
    return post["bbox"].decode_events({k: v.clone() for k, v in outputs.items()}, sizes, tags, 2, class_names=names)


def test_c2_b256_decoded_events_bf16_tier(c2_b256):
    """north_star: "decoded events identical".  The bf16 tier differs from the fp32 reference by ~5e-3 (pred_boxes by up
    to ~1e-2 absolute = 0.1 s on a 10 s clip), so a score within that distance of the 0.5 threshold (or a duration next to
    0.2 s, or two same-class events that nearly touch) may decode differently.  A clip is called STABLE when the oracle's own
    event list (labels and count) survives 8 random perturbations of its outputs of 3x the measured bf16 error; on stable
    clips the bf16 tier must give the same labels with onsets / offsets within 0.15 s (1.5e-2 of the clip, the pred_boxes
    bar of the bf16 tier) and scores within 0.03; over ALL clips the agreement rate is reported and must be >= 0.9."""
    post, ref, out = c2_b256["post"], c2_b256["ref"], c2_b256["out"]
    names = [f"class{i}" for i in range(10)]
    dev = {k: ref[k].cuda() for k in ("pred_logits", "pred_boxes", "at")}
    ev_ref = _decode(post, dev, names)
    ev_got = _decode(post, {k: out[k] for k in dev}, names)
    assert sum(len(e) for e in ev_ref) > 50
    g = torch.Generator(device="cuda").manual_seed(5)
    stable = np.ones(len(ev_ref), dtype=bool)
    for _ in range(8):
        pert = {"pred_logits": dev["pred_logits"] + 3 * 5e-3 * dev["pred_logits"].std() * torch.randn(dev["pred_logits"].shape, generator=g, device="cuda"),
                "pred_boxes": (dev["pred_boxes"] + 3 * 1e-3 * torch.randn(dev["pred_boxes"].shape, generator=g, device="cuda")).clamp(0, 1),
                "at": (dev["at"] + 3 * 3e-3 * torch.randn(dev["at"].shape, generator=g, device="cuda")).clamp(0, 1)}
        ev_p = _decode(post, pert, names)
        for i, (a, b) in enumerate(zip(ev_ref, ev_p)):
            if [e[0] for e in a] != [e[0] for e in b]:
                stable[i] = False
    agree, dt_max, ds_max = 0, 0.0, 0.0
    for i, (a, b) in enumerate(zip(ev_ref, ev_got)):
        same = [e[0] for e in a] == [e[0] for e in b]
        agree += int(same)
        if stable[i]:
            assert same, (i, a, b)
        if same:
            for ea, eb in zip(a, b):
                dt_max = max(dt_max, abs(ea[1] - eb[1]), abs(ea[2] - eb[2]))
                ds_max = max(ds_max, abs(ea[3] - eb[3]))
    rate = agree / len(ev_ref)
    print(f"decoded events, bf16 tier vs fp32 oracle: {agree}/{len(ev_ref)} clips identical labels ({rate:.3f}); "
          f"{int(stable.sum())} stable clips all identical; {sum(len(e) for e in ev_ref)} events; "
          f"max onset/offset difference {dt_max:.3f} s, max score difference {ds_max:.4f}")
    assert stable.sum() >= len(ev_ref) // 2
    assert rate >= 0.9
    assert dt_max < 0.15 and ds_max < 0.03


def test_c5_b200_spsedt_forward():
    """Config 5 at its full size: 200 clips + 2000 patches [1,128,64] through the eval branch (sedt/spsedt.py:70-75)."""
    args = spec.config_args("c5")
    args.precision = "bf16"
    sd = synth.synth_state_dict(args, 15)
    model, _, _ = build_model(args)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    B, P = 200, 10
    x = synth.synth_clips(B, 496, 64, seed=9)
    patches = synth.synth_patches(B, P, 128, 64, seed=9)
    mask = torch.zeros(B, 496, 64, dtype=torch.bool)
    with torch.no_grad():
        model((x[:2].cuda(), mask[:2].cuda()), patches[:2].cuda())
        k0 = _lib.kernel_kind_counts()
        out = model((x.cuda(), mask.cuda()), patches.cuda())
        torch.cuda.synchronize()
    kinds = _kinds_delta(k0)
    print("kernel kinds of one config-5 forward:", kinds)
    assert kinds["stem_tc"] == 2 and kinds["conv_tc3_2sm"] > 0
    got = _flat_outputs(out)
    idx = list(range(0, B, 8))                      # 25 clips = 250 patches through the fp32 oracle
    ref = _flat_outputs(sedt_oracle.spsedt_forward(sd, args, x[idx], mask[idx], patches[idx]))
    for k, r in ref.items():
        g = got[k]
        if k == "gt_feature":
            g = g.view(B, P, -1)[idx].reshape(-1, g.shape[-1])
        else:
            g = g[idx]
        e = rel_l2(g, r)
        print(f"c5 B=200 bf16 vs fp32 oracle {k}: rel-L2 {e:.2e}")
        assert e < 3e-2, (k, e)


def _loss(out, R, sl=slice(None)):
    tot = (out["pred_logits"] * R["l"][-1][sl]).sum() + (out["pred_boxes"] * R["b"][-1][sl]).sum()
    tot = tot + (out["at"] * R["a"][sl]).sum()
    for i, aux in enumerate(out["aux_outputs"]):
        tot = tot + (aux["pred_logits"] * R["l"][i][sl]).sum() + (aux["pred_boxes"] * R["b"][i][sl]).sum()
    return tot


def test_c4_training_step_matches_bf16_rounding_oracle():
    """Config 4's shape: E=6, Q=20, B=64 clips of 496 frames, one forward + backward through the native kernels (eager
    launches: the same kernels the graph replays) vs autograd through the bf16-rounding oracle.

    What the comparison can and cannot show.  Two bf16 implementations of this forward that round at the same points
    still differ by ~2e-3 (fp32 summation order -> bf16 ulp flips, amplified through ~60 layers; measured below), so a
    fraction ~2e-3 of the ReLU units sits on the other side of zero and each contributes a full-size error to the
    gradient: relative L2 ~ sqrt(flipped fraction) = 4-9 % for the deep layers (was 6-15 % against the fp32 oracle).
    That noise is (nearly) orthogonal to the true gradient, so a mis-scaled or mis-routed layer is caught by the
    PROJECTION of the kernel gradient on the oracle gradient, a = <g, r> / <r, r>, which the flips leave at 1:
        every trainable tensor: rel-L2 <= 0.10, cosine >= 0.995, |a - 1| <= 0.04 (tensors with >= 256 elements; 0.08 below; measured 0.02 / 0.05).
    The individual backward kernels are held to tight bars on identical inputs in tests/test_gpu_backward_ops.py."""
    args = spec.config_args("c2")
    args.precision, args.dropout = "bf16", 0.0
    sd = synth.synth_state_dict(args, 21)
    model, _, _ = build_model(args)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    B, D, Q, C = 64, args.dec_layers, args.num_queries, args.num_classes
    clips = synth.synth_clips(B, 496, 64, seed=30)
    g = torch.Generator().manual_seed(22)
    R = {"l": torch.randn(D, B, Q, C + 1, generator=g), "b": torch.randn(D, B, Q, 2, generator=g), "a": torch.randn(B, C, generator=g)}
    Rc = {k: v.cuda() for k, v in R.items()}
    k0 = _lib.kernel_kind_counts()
    out = model(clips.cuda())
    _loss(out, Rc).backward()
    torch.cuda.synchronize()
    kinds = _kinds_delta(k0)
    print("kernel kinds of one config-4 step:", kinds)
    assert kinds["wgrad_tc"] > 50 and kinds["attention_bwd_tc"] == args.enc_layers + 2 * D
    named = {n: p for n, p in model.named_parameters() if p.requires_grad}
    sdr = {k: v.clone().float() for k, v in sd.items()}
    for n in named:
        sdr[n].requires_grad_(True)
    fwd = {}
    for i in range(0, B, 16):                       # the loss is a sum over clips: gradients of the chunks add up
        sl = slice(i, i + 16)
        ref = bf16_oracle.sedt_forward_bf16(sdr, args, clips[sl], grad=True)
        _loss(ref, R, sl).backward()
        for k in ("pred_logits", "pred_boxes", "at"):
            fwd.setdefault(k, []).append(ref[k].detach())
    for k, v in fwd.items():
        e = rel_l2(out[k], torch.cat(v))
        print(f"c4 train forward {k}: rel-L2 {e:.2e}")
        assert e < 6e-3, (k, e)
    rows = []
    for n, p in named.items():
        gk, gr = p.grad.detach().float().cpu().flatten(), sdr[n].grad.flatten()
        if float(gk.norm()) == 0.0 and float(gr.norm()) == 0.0:         # e.g. decoder.layers.0.norm1.weight: tgt = 0 -> LN(0) = 0
            continue
        rel = ((gk - gr).norm() / gr.norm().clamp_min(1e-20)).item()
        cos = (torch.dot(gk, gr) / (gk.norm() * gr.norm()).clamp_min(1e-30)).item()
        proj = (torch.dot(gk, gr) / torch.dot(gr, gr).clamp_min(1e-30)).item()
        rows.append((rel, cos, proj, gk.numel(), n))
    rows.sort(reverse=True)
    print("c4 training step, worst gradients (rel-L2, cosine, projection):", [(n, round(r, 4), round(c, 5), round(a, 4)) for r, c, a, _, n in rows[:8]])
    big = [abs(a - 1) for _, _, a, k, _ in rows if k >= 256]
    print(f"c4 training step: median rel-L2 {sorted(r for r, *_ in rows)[len(rows) // 2]:.4f}, max |projection - 1| "
          f"{max(big):.4f} (>= 256 elements), {max(abs(a - 1) for _, _, a, _, _ in rows):.4f} (all)")
    bad = [(n, round(r, 4), round(c, 5), round(a, 4)) for r, c, a, k, n in rows
           if r > 0.10 or c < 0.995 or abs(a - 1) > (0.04 if k >= 256 else 0.08)]
    assert not bad, f"{len(bad)} of {len(rows)} gradients off: {bad[:12]}"
