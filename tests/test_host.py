"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the
native weight table equals the reference state_dict, and the host-side mirrors behave like the
reference's (no kernels are launched here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import sedt_oracle
from sound_event_detection_transformer_b200 import _lib, spec, synth
from sound_event_detection_transformer_b200.sedt import PostProcess, build_model
from sound_event_detection_transformer_b200.utils import NestedTensor, nested_tensor_from_tensor_list


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sedt_b200.h")).read()
    declared = set(re.findall(r"SEDT_API[^;(]*?\b(sedt_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.sedt_abi_version() == 1


@pytest.mark.parametrize("cfg", ["c1", "c2", "c5"])
def test_weight_table_is_the_reference_state_dict(cfg):
    args = spec.config_args(cfg)
    model, criterion, post = build_model(args)
    sd = synth.synth_state_dict(args, 0)
    assert list(model.state_dict().keys()) == list(sd.keys()) or set(model.state_dict()) == set(sd)
    model.load_state_dict(sd, strict=True)
    lib = _lib.load()
    c = _lib.SedtConfig(**model._native_config())
    h = C.c_void_p()
    _lib.check(lib.sedt_model_create(C.byref(c), C.byref(h)))
    n = lib.sedt_model_num_weights(h)
    names = [lib.sedt_model_weight_name(h, i).decode() for i in range(n)]
    assert sorted(names) == sorted(sd.keys())
    for i, nm in enumerate(names):
        assert lib.sedt_model_weight_numel(h, i) == sd[nm].numel(), nm
    assert lib.sedt_model_packed_bytes(h) > 0
    assert lib.sedt_workspace_bytes(h, 2, 496, 64, 10 if cfg == "c5" else 0, 128 if cfg == "c5" else 0) > 0
    lib.sedt_model_destroy(h)
    # freeze policy of sedt/backbone.py:60-62
    frozen = [k for k, p in model.named_parameters() if not p.requires_grad]
    if args.lr_backbone > 0:
        assert all(("conv1" in k and "layer" not in k) or "layer1" in k for k in frozen) and len(frozen) == 11
    else:
        assert all("backbone" in k for k in frozen) and len(frozen) == 55


def test_feature_shape_matches_oracle_geometry():
    lib = _lib.load()
    for T in (500, 496, 128, 333, 61):
        h, w = C.c_int(), C.c_int()
        _lib.check(lib.sedt_feature_shape(T, 64, 1, C.byref(h), C.byref(w)))
        assert [(h.value, w.value)] == spec.feature_hw(T, 64, True)[-1:]


def test_errors_are_reported_not_thrown():
    lib = _lib.load()
    bad = _lib.SedtConfig(enc_layers=3, dec_layers=3, num_queries=10, num_classes=10, hidden_dim=512, nheads=8,
                          dim_feedforward=2048, dec_at=1, pre_norm=1, dilation=1, precision=0)
    h = C.c_void_p()
    rc = lib.sedt_model_create(C.byref(bad), C.byref(h))
    assert rc == -1 and b"hidden_dim" in lib.sedt_last_error()
    with pytest.raises(_lib.SedtError):
        _lib.check(rc)


def test_model_without_cuda_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    args = spec.config_args("c1")
    model, _, _ = build_model(args)
    model.eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        with torch.no_grad():
            model(synth.synth_clips(1, 128, 64))


def test_nested_tensor_semantics():
    a, b = torch.randn(1, 500, 64), torch.randn(1, 333, 64)
    nt = nested_tensor_from_tensor_list([a, b])
    x, m = nt.decompose()
    rx, rm = sedt_oracle.nested([a, b], None)
    assert torch.equal(x, rx) and torch.equal(m, rm) and nt.unpadded is False
    nt2 = nested_tensor_from_tensor_list(torch.randn(3, 1, 64, 64))
    assert nt2.unpadded is True and not nt2.mask.any()
    with pytest.raises(ValueError):
        nested_tensor_from_tensor_list([torch.randn(5, 5)])
    assert isinstance(nt[0:1], NestedTensor)


def test_postprocess_has_no_cpu_path():
    out = {"pred_logits": torch.randn(2, 20, 11), "pred_boxes": torch.rand(2, 20, 2)}
    with pytest.raises(RuntimeError):
        PostProcess()(out, torch.full((2,), 10.0))


def test_new_entry_points_have_no_cpu_path():
    """FusedAdamW / clip_grad_norm_ / prepare_clips / pseudo_labels refuse CPU tensors instead of falling back."""
    from sound_event_detection_transformer_b200.optim import FusedAdamW, clip_grad_norm_
    from sound_event_detection_transformer_b200.prepare import prepare_clips
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        FusedAdamW([p]).step(max_norm=0.1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        clip_grad_norm_([p], 0.1)
    with pytest.raises(ValueError):
        FusedAdamW([p], betas=(0.3, 0.999))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            prepare_clips([torch.zeros(4, 64)], 8, device="cpu")
        with pytest.raises(RuntimeError, match="no CPU path"):
            PostProcess().pseudo_labels({"pred_logits": torch.zeros(1, 4, 11), "pred_boxes": torch.zeros(1, 4, 2)}, torch.ones(1), torch.ones(10))


def test_fused_criterion_gate_is_host_logic():
    """The fused set-criterion path is only taken for the supervised default recipe; everything else routes to the per-clip
    path (sedt/sedt.py:309-352 semantics) -- checked on the predicate alone, no kernels involved."""
    import numpy as np
    args = spec.config_args("c1")
    _, criterion, _ = build_model(args)
    out = {"pred_logits": torch.zeros(2, 10, 11), "pred_boxes": torch.zeros(2, 10, 2)}
    tg = np.array([{"labels": torch.zeros(1, dtype=torch.int64), "boxes": torch.zeros(1, 2)}] * 2, dtype=object)
    assert not criterion._fused_ok(out, tg, slice(2), None, False, False)            # CPU tensors
    assert not criterion._fused_ok(out, tg, None, None, False, False)                # no strong clips
    assert not criterion._fused_ok(out, tg, slice(1, 2), None, False, False)         # strong set not at the front
    assert not criterion._fused_ok(out, tg, slice(2), None, True, False)             # fine_tune
    assert not criterion._fused_ok(out, tg, slice(2), None, False, True)             # normalize


def test_eval_lanes_host_logic():
    """lanes.py: argument checks are host logic; CPU models are refused (no CPU path)."""
    from sound_event_detection_transformer_b200.lanes import EvalLanes
    args = spec.config_args("c1")
    m, _, _ = build_model(args)
    m.eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        EvalLanes([m])
    with pytest.raises(ValueError):
        EvalLanes([m, m])
    with pytest.raises(ValueError):
        EvalLanes([])
    m.train()
    with pytest.raises(ValueError, match="eval"):
        EvalLanes([m])
