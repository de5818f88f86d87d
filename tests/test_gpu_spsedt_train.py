"""SP-SEDT pretraining step through the native kernels (sedt_forward_train_sp / sedt_backward_sp via SPSEDT.forward in
train() mode): the training branch of sedt/spsedt.py:63-69 (query drop, doubled query embedding, block-diagonal decoder
mask), the feature-reconstruction head and loss (sedt/sedt.py:263-282), and the gradients of every trainable parameter
(the backbone is frozen, train_spsedt.py:50).  Checked against golden outputs / gradients of the reference's own SPSEDT in
train() mode (mask injected at spsedt.py:65) and against autograd through the oracle.  GPU only.

Tolerances: forward rel-L2 <= 3e-2 (bf16 tier vs fp32 reference); gradients cosine >= 0.985, rel-L2 <= 0.16 (the same bars
as tests/test_gpu_train.py: bf16 forward -> a few flipped ReLU units, see that file's docstring)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import sedt_oracle
from sound_event_detection_transformer_b200 import spec, synth
from sound_event_detection_transformer_b200.sedt import build_model

pytestmark = pytest.mark.gpu
torch.set_num_threads(max(1, os.cpu_count() or 1))


def _model(args, seed):
    args.precision = "bf16"
    model, criterion, _ = build_model(args)
    model.load_state_dict(synth.synth_state_dict(args, seed), strict=True)
    return model.cuda().train(), criterion


def _case():
    fx = np.load(os.path.join(GOLDEN, "spsedt_train_c5_b2.npz"))
    B, T, seed = [int(v) for v in fx["meta"]]
    args = spec.config_args("c5"); args.dropout = 0.0; args.enc_layers = 2; args.dec_layers = 2
    x, patches = synth.synth_clips(B, T, 64, seed=seed), synth.synth_patches(B, 10, 128, 64, seed=seed)
    mask = torch.zeros(B, T, 64, dtype=torch.bool)
    return fx, args, seed, B, x, mask, patches


def test_spsedt_train_step_matches_reference_golden_and_oracle_autograd():
    fx, args, seed, B, x, mask, patches = _case()
    model, _ = _model(args, seed)
    assert not any(p.requires_grad for n, p in model.named_parameters() if n.startswith("backbone."))
    keep = torch.from_numpy(fx["keep"])
    out = model((x.cuda(), mask.cuda()), patches.cuda(), query_keep=keep.cuda())
    rel = lambda a, b: ((a.detach().float().cpu() - b).norm() / b.norm().clamp_min(1e-12)).item()
    assert rel(out["pred_logits"], torch.from_numpy(fx["pred_logits"])) < 3e-2
    assert rel(out["pred_boxes"], torch.from_numpy(fx["pred_boxes"])) < 3e-2
    assert rel(out["aux_outputs"][0]["pred_logits"], torch.from_numpy(fx["aux0_pred_logits"])) < 3e-2
    assert rel(out["pred_feature"].flatten()[::97], torch.from_numpy(fx["pred_feature"])) < 3e-2
    assert rel(out["gt_feature"].flatten()[::97], torch.from_numpy(fx["gt_feature"])) < 3e-2
    assert not out["gt_feature"].requires_grad and out["pred_feature"].requires_grad

    cpu_out = {k: (v if k != "aux_outputs" else v) for k, v in out.items()}
    Rsum = synth.sp_train_functional({k: (v.cpu() if torch.is_tensor(v) else [{a: b.cpu() for a, b in d.items()} for d in v])
                                      for k, v in cpu_out.items()}, B, 20, seed)
    Rsum.backward()
    torch.cuda.synchronize()
    named = {n: p for n, p in model.named_parameters() if p.requires_grad}
    # (1) the reference's own gradients (sampled fixture)
    for n in synth.SP_TRAIN_PARAMS:
        g = named[n].grad.detach().float().cpu()
        got = g.flatten()[::97] if g.numel() > 20000 else g.flatten()
        want = torch.from_numpy(fx["grad_" + n]).flatten()
        cos = (torch.dot(got, want) / (got.norm() * want.norm()).clamp_min(1e-30)).item()
        r = ((got - want).norm() / want.norm().clamp_min(1e-20)).item()
        assert cos > 0.985 and r < 0.16, (n, cos, r)
    # (2) every trainable tensor against autograd through the oracle
    sd = {k: v.clone() for k, v in synth.synth_state_dict(args, seed).items()}
    for n in named:
        sd[n].requires_grad_(True)
    with torch.enable_grad():
        ref = sedt_oracle.spsedt_forward.__wrapped__(sd, args, x, mask, patches, query_keep=keep.bool())
        synth.sp_train_functional(ref, B, 20, seed).backward()
    bad = []
    for n, p in named.items():
        g, r = p.grad.detach().float().cpu().flatten(), sd[n].grad.flatten()
        if float(g.norm()) == 0.0 and float(r.norm()) == 0.0:
            continue
        rl = ((g - r).norm() / r.norm().clamp_min(1e-20)).item()
        cos = (torch.dot(g, r) / (g.norm() * r.norm()).clamp_min(1e-30)).item()
        if rl > 0.16 or cos < 0.985:
            bad.append((n, round(rl, 4), round(cos, 5)))
    assert not bad, f"{len(bad)} of {len(named)} gradients off: {bad[:12]}"


def test_spsedt_train_mode_semantics():
    """train() draws the query-drop mask like the reference (torch.rand(Q, bs, 1) > mask_ratio), works under no_grad, needs
    num_patches patches, and the criterion's feature loss back-propagates through the native step."""
    fx, args, seed, B, x, mask, patches = _case()
    args.dropout = 0.1
    model, criterion = _model(args, seed)
    xs = (x.cuda(), mask.cuda())
    torch.manual_seed(5)
    want_keep = (torch.rand(20, B, 1, device="cuda") > model.mask_ratio)[:, :, 0].t().to(torch.uint8)
    torch.manual_seed(5)
    assert torch.equal(model.draw_query_keep(B, torch.device("cuda")), want_keep)
    with torch.no_grad():
        a = model(xs, patches.cuda())["pred_logits"].clone()
        b = model(xs, patches.cuda())["pred_logits"].clone()
    assert (a - b).abs().max() > 1e-4                       # fresh drop mask + dropout every call
    with pytest.raises(NotImplementedError, match="num_patches"):
        model(xs, patches[:, :5].cuda())
    # one optimisation step with the set criterion incl. loss_feature (engine.py:56-80 on the pretraining targets)
    targets = []
    for bi in range(B):
        c = torch.rand(10) * 0.8 + 0.1
        targets.append({"labels": torch.zeros(10, dtype=torch.long).cuda(), "boxes": torch.stack([c, torch.full((10,), 0.1)], -1).cuda(),
                        "patches": patches[bi].cuda(), "orig_size": torch.tensor(10.0).cuda()})
    out = model(xs, patches.cuda())
    losses, _ = criterion(out, np.array(targets, dtype=object), None, slice(B))
    assert "loss_feature" in losses and "loss_feature_0" in losses
    loss = sum(losses[k] * criterion.weight_dict[k] for k in losses if k in criterion.weight_dict)
    loss.backward()
    torch.cuda.synchronize()
    for n in ("feature_align.layers.1.weight", "patch2query.weight", "query_embed.weight", "input_proj.weight"):
        g = dict(model.named_parameters())[n].grad
        assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0, n
    model.eval()
    with torch.no_grad():
        e = model(xs, patches[:, :5].cuda())                 # the test branch still takes fewer patches
    assert e["pred_logits"].shape == (B, 10, 2)
