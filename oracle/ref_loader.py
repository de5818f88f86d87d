"""ORACLE (test infrastructure, not product code).

Imports the UNMODIFIED reference (oracle/_ref, see make_ref.py; /root/reference as a
fall-back in the build container) with the three harness-side shims of SURVEY.md
Appendix B, none of which touches a reference file:
  * torchvision.models.resnet50 is wrapped to force pretrained=False
    (sedt/backbone.py:98-100 hard-codes a weight download; there is no network);
  * a stub `dcase_util.data` module (utilities/BoxEncoder.py:4-5 imports two names it never uses);
  * SetCriterion is built directly when CUDA is absent (sedt/__init__.py:60 returns None then).
"""
from __future__ import annotations

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_loaded = None


def reference_root():
    for p in (os.path.join(HERE, "_ref"), os.environ.get("SEDT_REFERENCE_ROOT", "/root/reference")):
        if os.path.isdir(os.path.join(p, "sedt")) and os.path.exists(os.path.join(p, "config.py")):
            return p
    return None


def load():
    """Returns the reference's `sedt` package (build_model, build_matcher, ...) or raises RuntimeError."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not available: run `python oracle/make_ref.py` where /root/reference exists")
    import torchvision
    if not getattr(torchvision.models.resnet50, "_sedt_shim", False):
        r50 = torchvision.models.resnet50

        def resnet50(*a, **k):
            return r50(*a, **{**k, "pretrained": False})
        resnet50._sedt_shim = True
        torchvision.models.resnet50 = resnet50
    for name in ("dcase_util", "dcase_util.data"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["dcase_util.data"].DecisionEncoder = sys.modules["dcase_util.data"].ProbabilityEncoder = object
    if root not in sys.path:
        sys.path.insert(0, root)
    import sedt as ref_sedt                       # the reference package (top-level name `sedt`)
    if not os.path.abspath(ref_sedt.__file__).startswith(os.path.abspath(root)):
        raise RuntimeError(f"`import sedt` resolved to {ref_sedt.__file__}, not to the reference under {root}")
    _loaded = ref_sedt
    return ref_sedt


def build_reference_model(args, state_dict, device="cpu"):
    """The reference's own module (sedt.build_model(args)[0]) with `state_dict` loaded strictly, in eval mode."""
    ref = load()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model, _, _ = ref.build_model(args)
    missing = model.load_state_dict(state_dict, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.to(device).eval()


def build_reference_criterion(args, device="cpu"):
    """The reference's SetCriterion built as sedt/__init__.py:37-60 builds it (build_model returns None for it on a
    CUDA-less host because utilities/utils.py:85-110 has no else branch)."""
    ref = load()
    matcher = ref.build_matcher(args)
    weight_dict = {"loss_ce": args.ce_loss_coef, "loss_bbox": args.bbox_loss_coef, "loss_giou": args.giou_loss_coef}
    losses = ["labels", "boxes", "cardinality"]
    if not args.self_sup:
        if args.dec_at:
            weight_dict["loss_weak"] = args.weak_loss_coef
            losses += ["weak"]
    elif args.feature_recon:
        losses += ["feature"]
        weight_dict["loss_feature"] = 1
    if args.aux_loss:
        aux = {}
        for i in range(args.dec_layers - 1):
            aux.update({k + f"_{i}": v for k, v in weight_dict.items()})
        weight_dict.update(aux)
    crit = ref.SetCriterion(1 if args.self_sup else args.num_classes, matcher=matcher, weight_dict=weight_dict,
                            eos_coef=args.eos_coef, losses=losses)
    return crit.to(device)
