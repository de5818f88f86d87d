"""Generate tests/golden/*.npz|json by running the UNMODIFIED reference
(/root/reference) on seeded synthetic inputs and weights.

Run once in the build container (the reference cannot travel to the GPU box):
    python tests/golden/make_golden.py
The shims are harness-side only: force
pretrained=False (sedt/backbone.py:98-100 hard-codes a download), stub
dcase_util (utilities/BoxEncoder.py:4-5 imports it, never uses it), build
SetCriterion/matcher directly, and make Tensor.cuda the identity for
SPSEDT.forward (sedt/spsedt.py:37-38 hard-codes .cuda()).

Fixtures hold outputs in full (small) and strided samples of intermediates.
"""
import json
import os
import sys
import types
import warnings

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

_r50 = torchvision.models.resnet50
torchvision.models.resnet50 = lambda *a, **k: _r50(*a, **{**k, "pretrained": False})
for name in ("dcase_util", "dcase_util.data"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["dcase_util.data"].DecisionEncoder = sys.modules["dcase_util.data"].ProbabilityEncoder = object

from sedt import build_model, build_matcher  # noqa: E402  (the reference package)
from sedt.sedt import PostProcess  # noqa: E402
from utilities.BoxEncoder import BoxEncoder  # noqa: E402

from sound_event_detection_transformer_b200 import spec, synth  # noqa: E402

CLASSES = [f"class{i}" for i in range(10)]
SAMPLE = 97


def sample(t):
    return t.detach().flatten()[::SAMPLE].contiguous().numpy()


def build_ref(args, seed):
    model, _, _ = build_model(args)
    sd = synth.synth_state_dict(args, seed)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert list(model.state_dict().keys()) == list(sd.keys()) or set(model.state_dict()) == set(sd)
    model.eval()
    return model


def run_sedt(tag, args, clips, seed):
    model = build_ref(args, seed)
    taps = {}
    body = model.backbone[0].body
    hooks = [body.maxpool.register_forward_hook(lambda m, i, o: taps.__setitem__("stem", o))]
    for li in range(1, 5):
        hooks.append(getattr(body, f"layer{li}").register_forward_hook(
            lambda m, i, o, li=li: taps.__setitem__(f"layer{li}", o)))
    hooks.append(model.transformer.encoder.register_forward_hook(lambda m, i, o: taps.__setitem__("memory_sbc", o)))
    hooks.append(model.transformer.decoder.register_forward_hook(lambda m, i, o: taps.__setitem__("hs_dqbc", o)))
    with torch.no_grad():
        out = model(clips)
    for h in hooks:
        h.remove()
    fx = {"pred_logits": out["pred_logits"].numpy(), "pred_boxes": out["pred_boxes"].numpy()}
    if "at" in out:
        fx["at"] = out["at"].numpy()
    for i, aux in enumerate(out.get("aux_outputs", [])):
        fx[f"aux{i}_pred_logits"] = aux["pred_logits"].numpy()
        fx[f"aux{i}_pred_boxes"] = aux["pred_boxes"].numpy()
    for k in ("stem", "layer1", "layer2", "layer3", "layer4"):
        fx["tap_" + k] = sample(taps[k])
    fx["tap_memory"] = sample(taps["memory_sbc"].permute(1, 0, 2))          # [B,S,C]
    fx["tap_hs"] = sample(taps["hs_dqbc"].transpose(1, 2))                 # [D,B,Q,C]
    np.savez_compressed(os.path.join(HERE, f"sedt_{tag}.npz"), **fx)

    # decoded events through the reference's own PostProcess + BoxEncoder (engine.py:264-291)
    if "at" in out:
        B = out["pred_logits"].shape[0]
        sizes = torch.full((B,), 10.0)
        tags = (out["at"].reshape(B, -1) > 0.5).long()
        enc = BoxEncoder(list(CLASSES), seconds=10.0)
        events = {}
        for at_m in (1, 2, 3):
            res = PostProcess()({k: v.clone() for k, v in out.items() if k != "aux_outputs"}, sizes, tags, at_m)
            per_clip = []
            for r in res:
                r = {k: v.numpy() for k, v in r.items()}
                per_clip.append([[e[0], float(e[1]), float(e[2]), float(e[3])] for e in enc.decode_strong(r, 0.5)])
            events[str(at_m)] = per_clip
        with open(os.path.join(HERE, f"events_{tag}.json"), "w") as f:
            json.dump(events, f)
        n_ev = sum(len(c) for c in events["2"])
        print(f"  events at_m=2: {n_ev}")
    print(f"sedt_{tag}: logits std {out['pred_logits'].std():.4f} boxes std {out['pred_boxes'].std():.4f}")


def run_spsedt(tag, args, clips, patches, seed):
    torch.Tensor.cuda = lambda self, *a, **k: self          # spsedt.py:37-38
    model = build_ref(args, seed)
    mask = torch.zeros(clips.shape[0], clips.shape[2], clips.shape[3], dtype=torch.bool)
    with torch.no_grad():
        out = model((clips, mask), patches)
    fx = {"pred_logits": out["pred_logits"].numpy(), "pred_boxes": out["pred_boxes"].numpy(),
          "pred_feature": sample(out["pred_feature"]), "gt_feature": sample(out["gt_feature"])}
    for i, aux in enumerate(out.get("aux_outputs", [])):
        fx[f"aux{i}_pred_logits"] = aux["pred_logits"].numpy()
        fx[f"aux{i}_pred_boxes"] = aux["pred_boxes"].numpy()
    np.savez_compressed(os.path.join(HERE, f"spsedt_{tag}.npz"), **fx)
    print(f"spsedt_{tag}: logits std {out['pred_logits'].std():.4f}")


def run_matcher(tag, B, Q, C, kmin, kmax, seed, chunk=32, normalize=False):
    args = spec.default_args()
    matcher = build_matcher(args)
    outputs, targets = synth.synth_matcher_case(B, Q, C, kmin, kmax, seed)
    rows, cols, counts = [], [], []
    for s in range(0, B, chunk):
        o = {k: v[s:s + chunk] for k, v in outputs.items()}
        idx, coef = matcher(o, targets[s:s + chunk], normalize=normalize)
        for (r, c), cf in zip(idx, coef):
            rows.append(r.numpy()); cols.append(c.numpy()); counts.append(len(r))
            assert cf.shape[0] == len(r)
    np.savez_compressed(os.path.join(HERE, f"matcher_{tag}.npz"),
                        rows=np.concatenate(rows) if rows else np.zeros(0, np.int64),
                        cols=np.concatenate(cols) if cols else np.zeros(0, np.int64),
                        counts=np.asarray(counts, np.int32),
                        meta=np.asarray([B, Q, C, kmin, kmax, seed], np.int64))
    print(f"matcher_{tag}: {B} clips, {int(np.sum(counts))} pairs")


def run_matcher_variant(tag, B, Q, C, kmin, kmax, seed, fine_tune=False, normalize=False, fl=False, epsilon=1.0, alpha=1.0,
                        rng_seed=1234):
    """The matcher's other branches (sedt/matcher.py:77-82 focal cost, :99-121 fine_tune relaxation, :123-133 coefficients).
    fine_tune draws torch.rand per clip from the global CPU generator: it is seeded with rng_seed right before the call and
    the test does the same."""
    args = spec.default_args()
    args.epsilon, args.alpha = epsilon, alpha
    matcher = build_matcher(args)
    outputs, targets = synth.synth_matcher_case(B, Q, C, kmin, kmax, seed)
    torch.manual_seed(rng_seed)
    idx, coef = matcher(outputs, targets, fine_tune=fine_tune, normalize=normalize, fl=fl)
    np.savez_compressed(os.path.join(HERE, f"matcher_{tag}.npz"),
                        rows=np.concatenate([r.numpy() for r, _ in idx]), cols=np.concatenate([c.numpy() for _, c in idx]),
                        coef=np.concatenate([c.numpy() for c in coef]).astype(np.float32),
                        counts=np.asarray([len(r) for r, _ in idx], np.int32),
                        meta=np.asarray([B, Q, C, kmin, kmax, seed, int(fine_tune), int(normalize), int(fl), rng_seed], np.int64),
                        fmeta=np.asarray([epsilon, alpha], np.float64))
    print(f"matcher_{tag}: {B} clips, {sum(len(r) for r, _ in idx)} pairs")


def synth_teacher_case(B, Q, C, seed):
    """Teacher-shaped outputs for the pseudo-label path: a few confident queries per clip, clustered intervals."""
    g = torch.Generator().manual_seed(9000 + seed)
    logits = torch.randn(B, Q, C + 1, generator=g)
    hot = torch.randint(0, 4, (B, Q), generator=g)
    logits.scatter_add_(2, hot.unsqueeze(-1), ((torch.rand(B, Q, generator=g) > 0.4).float() * 6.0).unsqueeze(-1))
    boxes = torch.stack([torch.rand(B, Q, generator=g), torch.rand(B, Q, generator=g) * 0.3], -1)
    boxes[:, ::5, 1] *= 0.05
    at = torch.rand(B, C, generator=g)
    return logits, boxes, at


def run_pseudo_labels(tag, B, Q, C, seed, del_overlap=True):
    """engine.get_pseudo_labels (engine.py:300-348) of the reference on seeded teacher outputs.  engine.py imports the data /
    metric stack (librosa, sed_eval, psds_eval), absent here and never used by this function: stubbed like dcase_util."""
    for name in ("librosa", "soundfile", "sed_eval", "psds_eval", "sed_eval.sound_event", "sed_eval.util", "librosa.display",
                 "librosa.feature"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["psds_eval"].PSDSEval = object
    sys.modules["psds_eval"].plot_psd_roc = None
    import collections
    import engine
    logits, boxes, at = synth_teacher_case(B, Q, C, seed)
    thr = torch.linspace(0.45, 0.7, C)
    sizes = torch.full((B,), 10.0)
    targets = [dict() for _ in range(B)]
    counter = collections.Counter()
    out = engine.get_pseudo_labels({"pred_logits": logits, "pred_boxes": boxes, "at": at}, {"bbox": PostProcess()}, sizes, targets,
                                   counter, del_overlap=del_overlap, classwise_threshold=thr)
    np.savez_compressed(os.path.join(HERE, f"pseudo_{tag}.npz"),
                        labels=np.concatenate([t["labels"].numpy() for t in out]).astype(np.int64),
                        boxes=np.concatenate([t["boxes"].numpy().reshape(-1, 2) for t in out]).astype(np.float32),
                        counts=np.asarray([len(t["labels"]) for t in out], np.int32),
                        meta=np.asarray([B, Q, C, seed, int(del_overlap)], np.int64), thr=thr.numpy())
    print(f"pseudo_{tag}: {B} clips, {sum(len(t['labels']) for t in out)} pseudo events")


def synth_decode_cases(n, Q, seed):
    """PostProcess-shaped results with few classes and heavily overlapping intervals (chains of same-class overlaps)."""
    rng = np.random.default_rng(9700 + seed)
    cases = []
    for _ in range(n):
        onset = rng.uniform(0.0, 8.0, Q).astype(np.float32)
        dur = rng.uniform(0.05, 3.0, Q).astype(np.float32)
        cases.append({"scores": rng.uniform(0.3, 1.0, Q).astype(np.float32), "labels": rng.integers(0, 3, Q).astype(np.int64),
                      "boxes": np.stack([onset, onset + dur], -1).astype(np.float32)})
    return cases


def run_decode_chains(tag, n, Q, seed):
    """BoxEncoder.decode_strong (utilities/BoxEncoder.py:179-226) of the reference on overlap-heavy results: pins the order of the
    suppression loop (delete the weaker neighbour, re-compare) that model outputs rarely exercise."""
    enc = BoxEncoder(list(CLASSES), seconds=10.0)
    out = []
    for r in synth_decode_cases(n, Q, seed):
        out.append([[e[0], float(e[1]), float(e[2]), float(e[3])] for e in enc.decode_strong(r, 0.5)])
    with open(os.path.join(HERE, f"decode_{tag}.json"), "w") as f:
        json.dump({"meta": [n, Q, seed], "events": out}, f)
    print(f"decode_{tag}: {n} clips, {sum(len(c) for c in out)} events")


def synth_db_clips(lengths, F, seed):
    g = torch.Generator().manual_seed(9500 + seed)
    return [(torch.randn(t, F, generator=g) * 12.0 - 40.0).numpy().astype(np.float32) for t in lengths]


def run_prepare(tag, lengths, frames, seed):
    """The reference's own PadOrTrunc -> ToTensor -> Normalize (utilities/BoxTransforms.py:91-238, utilities/Scaler.py) on
    seeded dB-domain features of ragged lengths.  ApplyLog is librosa (absent here) and is not part of this fixture."""
    for name in ("librosa", "soundfile", "librosa.display", "librosa.feature"):
        sys.modules.setdefault(name, types.ModuleType(name))
    from utilities.BoxTransforms import Compose, Normalize, PadOrTrunc, ToTensor
    from utilities.Scaler import Scaler
    F = 64
    g = torch.Generator().manual_seed(9600 + seed)
    sc = Scaler()
    sc.mean_ = (torch.randn(F, generator=g, dtype=torch.float64) * 5.0 - 40.0).numpy()
    sc.std_ = (torch.rand(F, generator=g, dtype=torch.float64) * 10.0 + 5.0).numpy()
    tf = Compose([PadOrTrunc(nb_frames=frames), ToTensor(unsqueeze_axis=0), Normalize(scaler=sc)])
    outs = []
    for clip in synth_db_clips(lengths, F, seed):
        label = {"labels": np.zeros(0, np.int64), "boxes": np.zeros((0, 2), np.float32), "orig_size": np.asarray(10.0)}
        x, _ = tf((clip, label))
        outs.append(x.numpy())
    np.savez_compressed(os.path.join(HERE, f"prepare_{tag}.npz"), out=np.stack(outs).astype(np.float32), mean=sc.mean_, std=sc.std_,
                        lengths=np.asarray(lengths, np.int64), meta=np.asarray([frames, F, seed], np.int64))
    print(f"prepare_{tag}: {len(lengths)} clips -> {np.stack(outs).shape}")


def _stub_audio_modules():
    for name in ("librosa", "soundfile", "librosa.display", "librosa.feature"):
        sys.modules.setdefault(name, types.ModuleType(name))


def run_augment(tag, B, T, seed):
    """TimeMask -> FreqMask(fill_mode="mean") -> FreqShift exactly as get_transforms chains them (utilities/BoxTransforms.py:
    471-478), probabilities raised so that every transform fires often; np.random seeded per batch."""
    _stub_audio_modules()
    from utilities.BoxTransforms import FreqMask, FreqShift, TimeMask
    clips = synth.synth_db_clips([T] * B, 64, seed)
    tfs = [TimeMask(p=0.7), FreqMask(fill_mode="mean", p=0.7), FreqShift(p=0.7)]
    np.random.seed(4200 + seed)
    outs = []
    for c in clips:
        d = c.copy()
        for tf in tfs:
            d = tf.transform_data(d)
        outs.append(d)
    np.savez_compressed(os.path.join(HERE, f"augment_{tag}.npz"), out=np.stack(outs).astype(np.float32), meta=np.asarray([B, T, seed]))
    print(f"augment_{tag}: {np.stack(outs).shape}")


def run_query(tag, B, P, T, seed, fixed=False):
    """Query.transform_label (utilities/BoxTransforms.py:315-360) on normalised clips."""
    _stub_audio_modules()
    from utilities.BoxTransforms import Query
    x = synth.synth_clips(B, T, 64, seed=seed)
    boxes = synth.synth_patch_boxes(B, P, seed, fixed_len=(128 / T) if fixed else None)
    q = Query(fixed)
    outs = []
    for b in range(B):
        _, lab = q.transform_label((x[b], {"patches": 1, "boxes": boxes[b]}))
        outs.append(lab["patches"].numpy())
    np.savez_compressed(os.path.join(HERE, f"query_{tag}.npz"), out=np.stack(outs).astype(np.float32), meta=np.asarray([B, P, T, seed, int(fixed)]))
    print(f"query_{tag}: {np.stack(outs).shape}")


def run_mixup(tag, n_strong, n_weak, n_unl, T, seed, with_weak=True):
    """utilities/mixup.py: mixup_data on a [strong | weak | unlabelled] batch (np.random seeded)."""
    from utilities.mixup import mixup_data

    class NT:
        pass
    x, y = synth.synth_mixup_case(n_strong, n_weak, n_unl, T, 64, seed)
    nt = NT(); nt.tensors = x.clone()
    np.random.seed(4300 + seed)
    mask_weak = slice(n_strong, n_strong + n_weak) if with_weak else None
    xo, yo, s_sl, w_sl = mixup_data(nt, np.array(y, dtype=object), slice(n_strong), mask_weak, mix_up_ratio=0.5, max_events=20, alpha=3)
    fx = {"out": xo.tensors.numpy().astype(np.float32), "meta": np.asarray([n_strong, n_weak, n_unl, T, seed, int(with_weak)]),
          "slices": np.asarray([s_sl.start or 0, s_sl.stop, w_sl.start, w_sl.stop]),
          "n_labels": np.asarray([len(l["labels"]) for l in yo]), "n_boxes": np.asarray([len(l["boxes"]) for l in yo]),
          "labels": np.concatenate([l["labels"].numpy().reshape(-1) for l in yo] + [np.zeros(0, np.int64)]),
          "boxes": np.concatenate([l["boxes"].numpy().reshape(-1, 2) for l in yo if len(l["boxes"])] + [np.zeros((0, 2), np.float32)]),
          "ratio": np.concatenate([(l["ratio"].numpy() if "ratio" in l else -np.ones(len(l["labels"]), np.float32)).reshape(-1)
                                   for l in yo] + [np.zeros(0, np.float32)])}
    np.savez_compressed(os.path.join(HERE, f"mixup_{tag}.npz"), **fx)
    print(f"mixup_{tag}: {fx['out'].shape}, strong {s_sl}, weak {w_sl}, mixed rows {(fx['ratio'] >= 0).sum()}")



SP_TRAIN_PARAMS, sp_train_functional = synth.SP_TRAIN_PARAMS, synth.sp_train_functional


def run_spsedt_train(tag, B, T, seed, drop_ratio=0.3):
    """SPSEDT.forward in train() mode (sedt/spsedt.py:63-69) with dropout 0 and an injected query-drop mask (torch.rand is
    patched for the one call at :65), outputs + gradients of a fixed functional w.r.t. a sample of the trainable parameters."""
    torch.Tensor.cuda = lambda self, *a, **k: self          # spsedt.py:37-38
    args = spec.config_args("c5"); args.dropout = 0.0; args.enc_layers = 2; args.dec_layers = 2
    model = build_ref(args, seed)
    model.train()
    x = synth.synth_clips(B, T, 64, seed=seed)
    patches = synth.synth_patches(B, 10, 128, 64, seed=seed)
    mask = torch.zeros(B, T, 64, dtype=torch.bool)
    keep = torch.rand(20, B, 1, generator=torch.Generator().manual_seed(8700 + seed)) > drop_ratio
    real_rand = torch.rand
    torch.rand = lambda *a, **k: torch.where(keep, torch.ones(1), torch.zeros(1))     # > mask_ratio exactly where keep
    try:
        out = model((x, mask), patches)
    finally:
        torch.rand = real_rand
    sp_train_functional(out, B, 20, seed).backward()
    named = dict(model.named_parameters())
    fx = {"keep": keep[:, :, 0].t().numpy().astype(np.uint8), "meta": np.asarray([B, T, seed]),
          "pred_logits": out["pred_logits"].detach().numpy(), "pred_boxes": out["pred_boxes"].detach().numpy(),
          "pred_feature": sample(out["pred_feature"]), "gt_feature": sample(out["gt_feature"]),
          "aux0_pred_logits": out["aux_outputs"][0]["pred_logits"].detach().numpy()}
    for n in SP_TRAIN_PARAMS:
        fx["grad_" + n] = sample(named[n].grad) if named[n].grad.numel() > 20000 else named[n].grad.numpy()
    assert all(not p.requires_grad for n, p in named.items() if n.startswith("backbone."))
    np.savez_compressed(os.path.join(HERE, f"spsedt_train_{tag}.npz"), **fx)
    print(f"spsedt_train_{tag}: kept {int(keep.sum())} of {keep.numel()} queries")



def run_criterion(tag, args, B, seed, kmin=0, kmax=10, fine_tune=False, normalize=False, fl=False, rng_seed=1234):
    """Reference SetCriterion (sedt/sedt.py:134-352) built directly (build_model returns None for it
    without CUDA) on seeded model-shaped outputs: every loss value and the gradient of the weighted sum."""
    from sedt.sedt import SetCriterion
    weight_dict = {"loss_ce": args.ce_loss_coef, "loss_bbox": args.bbox_loss_coef, "loss_giou": args.giou_loss_coef,
                   "loss_weak": args.weak_loss_coef}
    for i in range(args.dec_layers - 1):
        weight_dict.update({k + f"_{i}": v for k, v in list(weight_dict.items()) if "_" in k and not k[-1].isdigit()})
    crit = SetCriterion(args.num_classes, matcher=build_matcher(args), weight_dict=weight_dict, eos_coef=args.eos_coef,
                        losses=["labels", "boxes", "cardinality", "weak"])
    outputs, targets = synth.synth_criterion_case(B, args.num_queries, args.num_classes, args.dec_layers, kmin, kmax, seed)
    leaves = [outputs["pred_logits"], outputs["pred_boxes"], outputs["at"]]
    for a in outputs["aux_outputs"]:
        leaves += [a["pred_logits"], a["pred_boxes"]]
    for t in leaves:
        t.requires_grad_(True)
    torch.manual_seed(rng_seed)
    losses, indices = crit(outputs, np.array(targets, dtype=object), None, slice(B), fine_tune=fine_tune, normalize=normalize,
                           fl=fl)
    total = sum(losses[k] * weight_dict[k] for k in losses if k in weight_dict)
    total.backward()
    fx = {"loss_names": np.array(sorted(losses)), "loss_values": np.array([float(losses[k]) for k in sorted(losses)], np.float64),
          "total": np.float64(float(total)), "meta": np.asarray([B, kmin, kmax, seed], np.int64),
          "flags": np.asarray([int(fine_tune), int(normalize), int(fl), rng_seed], np.int64)}
    for i, t in enumerate(leaves):
        fx[f"grad_{i}"] = t.grad.numpy()
    np.savez_compressed(os.path.join(HERE, f"criterion_{tag}.npz"), **fx)
    print(f"criterion_{tag}: total {float(total):.5f}, {len(losses)} entries")


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if "--only-decode" in sys.argv:
        run_decode_chains("chains", 24, 20, seed=51)
        sys.exit(0)
    if "--only-sptrain" in sys.argv:
        run_spsedt_train("c5_b2", 2, 200, seed=15)
        sys.exit(0)
    if "--only-augment" in sys.argv:
        run_augment("b12", 12, 96, seed=51)
        run_query("p6", 3, 6, 200, seed=52)
        run_query("fixed", 2, 3, 300, seed=53, fixed=True)
        run_mixup("ss", 12, 6, 6, 40, seed=54)
        run_mixup("strong_only", 10, 0, 0, 40, seed=55, with_weak=False)
        run_mixup("weak_mix", 4, 10, 6, 24, seed=56)
        sys.exit(0)
    if "--only-prepare" in sys.argv:
        run_prepare("ragged", [40, 64, 90, 1, 63, 65], 64, seed=41)
        sys.exit(0)
    if "--only-pseudo" in sys.argv:
        run_pseudo_labels("q20", 48, 20, 10, seed=31)
        run_pseudo_labels("q10_keepall", 16, 10, 10, seed=32, del_overlap=False)
        sys.exit(0)
    if "--only-variants" in sys.argv:
        run_matcher_variant("v_fl", 64, 20, 10, 0, 10, seed=21, fl=True)
        run_matcher_variant("v_finetune", 64, 20, 10, 1, 10, seed=22, fine_tune=True, normalize=True, epsilon=1.0, alpha=1.0)
        run_matcher_variant("v_finetune_q10", 32, 10, 10, 1, 6, seed=23, fine_tune=True, epsilon=0.5, alpha=2.0, fl=True)
        run_criterion("v_fl", spec.config_args("c1"), 12, seed=24, fl=True)
        run_criterion("v_finetune", spec.config_args("c1"), 12, seed=25, kmin=1, kmax=6, fine_tune=True, normalize=True)
        sys.exit(0)
    if "--only-criterion" in sys.argv:
        run_criterion("c2", spec.config_args("c2"), 48, seed=1)
        run_criterion("c1_edges", spec.config_args("c1"), 16, seed=2, kmin=8, kmax=14)
        sys.exit(0)
    # config-1 shape (URBAN-SED: T=500, E=3, Q=10) and config-2 shape (DCASE: T=496, E=6, Q=20), batch 2
    run_sedt("c1_b2", spec.config_args("c1"), synth.synth_clips(2, 500, 64, seed=1), seed=11)
    run_sedt("c2_b2", spec.config_args("c2"), synth.synth_clips(2, 496, 64, seed=2), seed=12)
    # ragged list input -> real padding mask (utilities/utils.py:470-492), B=1 squeeze quirk (sedt.py:92)
    rag = [synth.synth_clips(1, 500, 64, seed=3)[0], synth.synth_clips(1, 333, 64, seed=4)[0],
           synth.synth_clips(1, 420, 64, seed=5)[0]]
    run_sedt("c1_ragged", spec.config_args("c1"), rag, seed=11)
    run_sedt("c1_b1", spec.config_args("c1"), synth.synth_clips(1, 500, 64, seed=6), seed=11)
    # no audio query / no aux variants of the head wiring (sedt.py:107-123)
    a = spec.config_args("c1"); a.dec_at = False; a.aux_loss = False
    run_sedt("c1_plain", a, synth.synth_clips(2, 256, 64, seed=7), seed=13)
    # post-norm transformer (--pre_norm flag flips it, train_sedt.py:98)
    a = spec.config_args("c1"); a.pre_norm = False
    run_sedt("c1_postnorm", a, synth.synth_clips(2, 256, 64, seed=8), seed=14)
    # SP-SEDT (config 5 shape, batch 2, 10 patches)
    run_spsedt("c5_b2", spec.config_args("c5"), synth.synth_clips(2, 496, 64, seed=9),
               synth.synth_patches(2, 10, 128, 64, seed=9), seed=15)
    # matcher: config-3 distribution, plus K=0 and K>Q edge set
    run_matcher("c3_small", 256, 20, 10, 0, 10, seed=3)
    run_matcher("c3_edges", 64, 20, 10, 18, 28, seed=4)
    run_matcher("urban_q10", 64, 10, 10, 0, 12, seed=5)
    run_matcher("c3_normalize", 32, 20, 10, 0, 10, seed=6, normalize=True)
    run_criterion("c2", spec.config_args("c2"), 48, seed=1)
    run_criterion("c1_edges", spec.config_args("c1"), 16, seed=2, kmin=8, kmax=14)
    run_augment("b12", 12, 96, seed=51)
    run_query("p6", 3, 6, 200, seed=52)
    run_query("fixed", 2, 3, 300, seed=53, fixed=True)
    run_mixup("ss", 12, 6, 6, 40, seed=54)
    run_mixup("strong_only", 10, 0, 0, 40, seed=55, with_weak=False)
    run_mixup("weak_mix", 4, 10, 6, 24, seed=56)
    run_spsedt_train("c5_b2", 2, 200, seed=15)
