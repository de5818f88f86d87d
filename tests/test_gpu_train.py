"""Training step through the native forward_train / backward kernels (bf16 tier) against autograd through the
fp32 oracle on the same seeded weights and clips.  GPU only.

Tolerance.  The forward runs in bf16 and agrees with the fp32 reference to ~0.5 % (outputs); a ReLU unit whose
pre-activation lies within that error of zero takes the other branch, and every flipped unit contributes a
full-size error to the gradient that passes through it, so the relative L2 error of a gradient grows like
sqrt(fraction of flipped units) with depth: measured 0.2-1 % at the heads, 3-4 % through the transformer,
6-15 % at layer3 / layer2 / conv0, with cosine similarity >= 0.989 and norms within 5 % everywhere (any
bf16 training step compared with fp32 autograd shows this).  The bars: cosine >= 0.985 and relative L2
<= 0.16 for every parameter, relative L2 <= 2e-2 for the parameters above the first ReLU (class_embed,
bbox_embed.layers.2, weak_class_embed, decoder.norm).  The single backward kernels are held to tight
tolerances against torch on identical inputs in tests/test_gpu_backward_ops.py.
"""
import os

import pytest
import torch

from oracle import sedt_oracle
from sound_event_detection_transformer_b200 import spec, synth
from sound_event_detection_transformer_b200.sedt import build_model

pytestmark = pytest.mark.gpu
torch.set_num_threads(max(1, os.cpu_count() or 1))


def _loss(out, R):
    """A fixed random linear functional of every output (all decoder layers), so that every path gets a gradient."""
    tot = (out["pred_logits"] * R["l"][-1]).sum() + (out["pred_boxes"] * R["b"][-1]).sum()
    if "at" in out:
        tot = tot + (out["at"] * R["a"]).sum()
    for i, aux in enumerate(out.get("aux_outputs", [])):
        tot = tot + (aux["pred_logits"] * R["l"][i]).sum() + (aux["pred_boxes"] * R["b"][i]).sum()
    return tot


def _setup(args, seed, B, T, masked=False):
    args.dropout = 0.0
    args.precision = "bf16"
    sd = synth.synth_state_dict(args, seed)
    model, _, _ = build_model(args)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    if masked:
        clips = [synth.synth_clips(1, T, 64, seed=31)[0], synth.synth_clips(1, T - 37, 64, seed=32)[0]][:B]
    else:
        clips = synth.synth_clips(B, T, 64, seed=30)
    g = torch.Generator().manual_seed(seed + 1)
    D, Q, C = args.dec_layers, args.num_queries, args.num_classes
    R = {"l": torch.randn(D, B, Q, C + 1, generator=g), "b": torch.randn(D, B, Q, 2, generator=g),
         "a": torch.randn(B, C, generator=g)}
    return sd, model, clips, R


def _reference_grads(sd, args, clips, R, names):
    sdr = {k: v.clone().float() for k, v in sd.items()}
    for n in names:
        sdr[n].requires_grad_(True)
    with torch.enable_grad():
        ref = sedt_oracle.sedt_forward.__wrapped__(sdr, args, clips)
        _loss(ref, R).backward()
    return {n: sdr[n].grad for n in names}, ref


@pytest.mark.parametrize("masked", [False, True])
def test_training_step_gradients_match_reference_autograd(masked):
    args = spec.config_args("c1")
    args.enc_layers, args.dec_layers = 2, 2
    B, T = 2, 200
    sd, model, clips, R = _setup(args, 21, B, T, masked)
    xin = [c.cuda() for c in clips] if isinstance(clips, list) else clips.cuda()
    out = model(xin)
    Rc = {k: v.cuda() for k, v in R.items()}
    _loss(out, Rc).backward()
    torch.cuda.synchronize()
    named = {n: p for n, p in model.named_parameters() if p.requires_grad}
    assert "backbone.0.body.conv0.weight" in named and "backbone.0.body.layer1.0.conv1.weight" not in named
    ref_grads, ref = _reference_grads(sd, args, clips, R, list(named))
    # forward outputs of the train-mode path
    for k in ("pred_logits", "pred_boxes", "at"):
        assert ((out[k].detach().cpu() - ref[k]).norm() / ref[k].norm()).item() < 3e-2, k
    worst = []
    for n, p in named.items():
        assert p.grad is not None, n
        g, r = p.grad.detach().float().cpu().flatten(), ref_grads[n].flatten()
        rel = ((g - r).norm() / r.norm().clamp_min(1e-20)).item()
        cos = (torch.dot(g, r) / (g.norm() * r.norm()).clamp_min(1e-30)).item()
        worst.append((rel, cos, n))
    worst.sort(reverse=True)
    bad = [(n, round(rel, 4), round(cos, 5)) for rel, cos, n in worst
           if (rel > 0.16 or cos < 0.985) and not (rel == 0.0 and cos == 0.0)]          # 0/0: gradient exactly zero in both
    assert not bad, f"{len(bad)} of {len(worst)} gradients off: {bad[:12]}"
    top = ("class_embed.", "bbox_embed.layers.2.", "weak_class_embed.", "transformer.decoder.norm.")
    for rel, cos, n in worst:
        if n.startswith(top):
            assert rel < 2e-2, (n, rel)


def test_training_step_frozen_backbone_and_repeatable():
    """lr_backbone = 0 freezes the whole backbone (sedt/backbone.py:135-141): only transformer / head gradients,
    and two identical steps give the same gradients up to the atomics' summation order."""
    args = spec.config_args("c1")
    args.enc_layers, args.dec_layers = 1, 2
    args.lr_backbone = 0.0
    sd, model, clips, R = _setup(args, 22, 2, 160)
    Rc = {k: v.cuda() for k, v in R.items()}
    assert not any(p.requires_grad for n, p in model.named_parameters() if n.startswith("backbone."))
    grads = []
    for _ in range(2):
        model.zero_grad(set_to_none=True)
        _loss(model(clips.cuda()), Rc).backward()
        torch.cuda.synchronize()
        grads.append({n: p.grad.clone() for n, p in model.named_parameters() if p.requires_grad})
    for n in grads[0]:
        a, b = grads[0][n].float(), grads[1][n].float()
        assert ((a - b).norm() / a.norm().clamp_min(1e-20)).item() < 1e-4, n
    ref_grads, _ = _reference_grads(sd, args, clips, R, list(grads[0]))
    for n, g in grads[0].items():
        r = ref_grads[n].flatten()
        rel = ((g.float().cpu().flatten() - r).norm() / r.norm().clamp_min(1e-20)).item()
        assert rel < 0.16, (n, rel)


def test_training_mode_requirements():
    args = spec.config_args("c1")
    args.precision = "bf16"
    args.pre_norm = False
    model, _, _ = build_model(args)
    model.load_state_dict(synth.synth_state_dict(args, 3))
    model.cuda().train()
    with pytest.raises(NotImplementedError, match="pre-norm"):
        model(synth.synth_clips(1, 128, 64).cuda())


# ---- dropout ---------------------------------------------------------------------------------------------------
def _mask_provider(args, B, S, seed, step, p):
    """oracle.DROPOUT hook that applies exactly the masks the training kernels draw (sedt_op_dropout_mask)."""
    import ctypes as C
    import re
    from sound_event_detection_transformer_b200 import _lib
    lib = _lib.load()
    H = args.nheads

    def flags(site, n):
        out = torch.empty(n, dtype=torch.uint8, device="cuda")
        _lib.check(lib.sedt_op_dropout_mask(out.data_ptr(), n, C.c_uint64(seed), C.c_uint64(step), site, p, _lib.current_stream()))
        return out.cpu().float()

    enc_k = {"attn": 0, "drop1": 1, "hidden": 2, "drop2": 3}
    dec_k = {"self_attn": 0, "drop1": 1, "cross_attn": 2, "drop2": 3, "hidden": 4, "drop3": 5}

    def hook(site, t):
        m = re.match(r"transformer\.(encoder|decoder)\.layers\.(\d+)\.(\w+)$", site)
        kind, layer, what = m.group(1), int(m.group(2)), m.group(3)
        sid = 8 * layer + enc_k[what] if kind == "encoder" else 1024 + 8 * layer + dec_k[what]
        if what in ("attn", "self_attn", "cross_attn"):          # [B*H, L, Lk] <- ((b*H + h)*128 + i)*128 + j
            BH, L, Lk = t.shape
            keep = flags(sid, BH * 128 * 128).view(BH, 128, 128)[:, :L, :Lk]
        else:                                                     # sequence-first [L, B, C] <- row-major [B*L, C]
            L, Bn, Cc = t.shape
            keep = flags(sid, Bn * L * Cc).view(Bn, L, Cc).permute(1, 0, 2)
        return t * keep / (1.0 - p)
    return hook


def test_training_step_with_dropout_matches_reference_with_the_same_masks():
    """dropout 0.1 (the reference's default): the fp32 oracle is given the kernels' own Philox masks at every dropout
    site, so outputs and gradients must agree as in the dropout-free test."""
    args = spec.config_args("c1")
    args.enc_layers, args.dec_layers = 2, 2
    B, T, p = 2, 200, 0.1
    sd, model, clips, R = _setup(args, 23, B, T)
    model.transformer.dropout = p
    torch.manual_seed(1234)
    out = model(clips.cuda())
    Rc = {k: v.cuda() for k, v in R.items()}
    _loss(out, Rc).backward()
    torch.cuda.synchronize()
    rt = model._rt
    named = {n: pp for n, pp in model.named_parameters() if pp.requires_grad}
    S = out["pred_logits"].shape[0]
    sedt_oracle.DROPOUT = _mask_provider(args, B, S, rt._seed, 1, p)        # first step on a fresh tape: step counter = 1
    try:
        ref_grads, ref = _reference_grads(sd, args, clips, R, list(named))
    finally:
        sedt_oracle.DROPOUT = None
    for k in ("pred_logits", "pred_boxes", "at"):
        assert ((out[k].detach().cpu() - ref[k]).norm() / ref[k].norm()).item() < 3e-2, k
    bad = []
    for n, pp in named.items():
        g, r = pp.grad.detach().float().cpu().flatten(), ref_grads[n].flatten()
        if r.norm() == 0 and g.norm() == 0:
            continue
        rel = ((g - r).norm() / r.norm().clamp_min(1e-20)).item()
        cos = (torch.dot(g, r) / (g.norm() * r.norm()).clamp_min(1e-30)).item()
        if rel > 0.16 or cos < 0.985:
            bad.append((n, round(rel, 4), round(cos, 5)))
    assert not bad, f"{len(bad)} gradients off: {bad[:12]}"
    # and the masks are really there: the dropout-free oracle must NOT match
    ref0 = sedt_oracle.sedt_forward(sd, args, clips)
    assert ((out["pred_logits"].detach().cpu() - ref0["pred_logits"]).norm() / ref0["pred_logits"].norm()).item() > 5e-2


def test_dropout_masks_change_every_step_and_eval_is_unaffected():
    args = spec.config_args("c1")
    args.enc_layers, args.dec_layers = 1, 1
    sd, model, clips, R = _setup(args, 24, 2, 160)
    model.transformer.dropout = 0.1
    x = clips.cuda()
    for graph in (False, True):
        model.use_cuda_graph = graph
        a = model(x)["pred_logits"].detach().clone()
        b = model(x)["pred_logits"].detach().clone()
        assert (a - b).abs().max() > 1e-3, graph                         # fresh masks each step (also under graph replay)
    model.eval()
    with torch.no_grad():
        e1 = model(x)["pred_logits"].clone()
        e2 = model(x)["pred_logits"].clone()
    assert torch.equal(e1, e2)


# ---- several forwards in flight / gradient accumulation -----------------------------------------------------
def _small_train_model(seed, dropout=0.0):
    args = spec.config_args("c1")
    args.enc_layers, args.dec_layers = 1, 1
    sd, model, _, _ = _setup(args, seed, 2, 160)
    model.transformer.dropout = dropout
    return args, sd, model


@pytest.mark.parametrize("graph", [False, True])
def test_two_forwards_then_one_backward(graph):
    """engine.py:134-170: model(labelled), model(unlabelled), ONE total.backward().  Each forward owns its tape: the summed
    gradient equals the sum of the two single-batch gradients."""
    args, sd, model = _small_train_model(25)
    model.use_cuda_graph = graph
    xa, xb = synth.synth_clips(2, 160, 64, seed=40).cuda(), synth.synth_clips(2, 160, 64, seed=41).cuda()
    g = torch.Generator().manual_seed(7)
    D, Q, C = args.dec_layers, args.num_queries, args.num_classes
    R = {"l": torch.randn(D, 2, Q, C + 1, generator=g).cuda(), "b": torch.randn(D, 2, Q, 2, generator=g).cuda(),
         "a": torch.randn(2, C, generator=g).cuda()}
    single = []
    for x in (xa, xb):
        model.zero_grad(set_to_none=True)
        _loss(model(x), R).backward()
        torch.cuda.synchronize()
        single.append({n: p.grad.clone() for n, p in model.named_parameters() if p.requires_grad})
    model.zero_grad(set_to_none=True)
    oa = model(xa)
    ob = model(xb)
    assert model._rt.train_slots_in_use() == 2
    (_loss(oa, R) + _loss(ob, R)).backward()
    torch.cuda.synchronize()
    assert model._rt.train_slots_in_use() == 0
    for n, p in model.named_parameters():
        if p.requires_grad:
            want = single[0][n] + single[1][n]
            assert ((p.grad - want).norm() / want.norm().clamp_min(1e-20)).item() < 2e-3, n     # fp32 atomics order only


def test_backward_twice_raises_instead_of_using_a_stale_tape():
    args, sd, model = _small_train_model(26)
    x = synth.synth_clips(2, 160, 64, seed=42).cuda()
    out = model(x)
    loss = out["pred_logits"].sum()
    loss.backward(retain_graph=True)
    model(x)["pred_logits"].sum().backward()          # reuses the slot
    with pytest.raises(RuntimeError, match="overwritten"):
        loss.backward()


def test_gradient_accumulation_in_graph_mode():
    """engine.py:75-80 (accumrating_gradient_steps): two micro-batches accumulated into p.grad, and zero_grad(set_to_none=False)
    between steps, must not alias the graph's static gradient buffer."""
    args, sd, model = _small_train_model(27)
    model.use_cuda_graph = True
    xa, xb = synth.synth_clips(2, 160, 64, seed=43).cuda(), synth.synth_clips(2, 160, 64, seed=44).cuda()
    single = []
    for x in (xa, xb):
        model.zero_grad(set_to_none=True)
        model(x)["pred_logits"].square().sum().backward()
        torch.cuda.synchronize()
        single.append({n: p.grad.clone() for n, p in model.named_parameters() if p.requires_grad})
    for set_to_none in (True, False):
        model.zero_grad(set_to_none=set_to_none)
        for x in (xa, xb):
            model(x)["pred_logits"].square().sum().backward()
        torch.cuda.synchronize()
        for n, p in model.named_parameters():
            if p.requires_grad:
                want = single[0][n] + single[1][n]
                assert ((p.grad - want).norm() / want.norm().clamp_min(1e-20)).item() < 2e-3, (set_to_none, n)


def test_no_grad_train_mode_keeps_dropout_active():
    """engine.py:146-147: the mean-teacher forward runs under no_grad with the module in train() mode."""
    args, sd, model = _small_train_model(28, dropout=0.1)
    x = synth.synth_clips(2, 160, 64, seed=45).cuda()
    with torch.no_grad():
        a = model(x)["pred_logits"].clone()
        b = model(x)["pred_logits"].clone()
        model.eval()
        e = model(x)["pred_logits"].clone()
    assert (a - b).abs().max() > 1e-3 and (a - e).abs().max() > 1e-3
    assert model._rt.train_slots_in_use() == 0
