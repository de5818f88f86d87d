"""Fused FFN kernel (csrc/ffn_fused.cu; sedt/transformer.py:202-203) through the C ABI against torch on the same bf16
operands: fp32 accumulation, hidden activation rounded to bf16 (as the unfused path stores it), fp32 bias / residual."""
import pytest
import torch

from sound_event_detection_transformer_b200 import _lib

pytestmark = pytest.mark.gpu


def _ref(x, w1, b1, w2, b2, res):
    h = torch.relu(x.float() @ w1.float().t() + b1).to(torch.bfloat16)
    return res + h.float() @ w2.float().t() + b2


@pytest.mark.parametrize("M,ff", [(128, 256), (248, 2048), (1000, 512), (31744, 2048), (5376, 2048)])
def test_ffn_fused_matches_torch(M, ff):
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + ff)
    x = torch.randn(M, 256, generator=g).cuda().bfloat16()
    w1 = (torch.randn(ff, 256, generator=g) / 16).cuda().bfloat16()
    w2 = (torch.randn(256, ff, generator=g) / (ff ** 0.5)).cuda().bfloat16()
    b1, b2 = torch.randn(ff, generator=g).cuda() * 0.1, torch.randn(256, generator=g).cuda() * 0.1
    res = torch.randn(M, 256, generator=g).cuda()
    out = torch.full((M, 256), float("nan"), device="cuda")
    _lib.check(lib.sedt_op_ffn(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), res.data_ptr(),
                               out.data_ptr(), M, ff, _lib.current_stream()))
    torch.cuda.synchronize()
    want = _ref(x, w1, b1, w2, b2, res)
    assert torch.isfinite(out).all()
    err = (out - want).abs().max().item()
    assert err <= 2e-2 * max(1.0, want.abs().max().item()), err          # bf16 hidden rounding flips a few ulps
    assert ((out - want).norm() / want.norm()).item() < 2e-3
    # in place (out == residual), as the model calls it
    buf = res.clone()
    _lib.check(lib.sedt_op_ffn(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), buf.data_ptr(),
                               buf.data_ptr(), M, ff, _lib.current_stream()))
    torch.cuda.synchronize()
    assert torch.equal(buf, out)


def test_model_forward_with_fused_ffn_matches_two_gemm_path(tmp_path):
    """The whole bf16 forward with the fused FFN forced on (it is chosen by itself only from 128 x SM-count rows) against
    the same forward with it forced off, each in a fresh process (the knob is read once): same rounding points, so the
    outputs agree far inside the bf16 tier's tolerance."""
    import os
    import subprocess
    import sys
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, torch, numpy as np; sys.path.insert(0, %r)\n"
        "from sound_event_detection_transformer_b200 import spec, synth\n"
        "from sound_event_detection_transformer_b200.sedt import build_model\n"
        "args = spec.config_args('c2'); model, _, _ = build_model(args)\n"
        "model.load_state_dict(synth.synth_state_dict(args, 12)); model = model.cuda().eval()\n"
        "with torch.no_grad(): out = model(synth.synth_clips(3, 496, 64, seed=2).cuda())\n"
        "np.savez(sys.argv[1], logits=out['pred_logits'].cpu().numpy(), boxes=out['pred_boxes'].cpu().numpy(), at=out['at'].cpu().numpy())\n"
    ) % root
    outs = {}
    for mode in ("0", "1"):
        fn = str(tmp_path / f"ffn{mode}.npz")
        env = dict(os.environ, SEDT_FFN_FUSED=mode)
        subprocess.run([sys.executable, "-c", code, fn], check=True, env=env, timeout=300)
        outs[mode] = np.load(fn)
    for k in ("logits", "boxes", "at"):
        a, b = outs["0"][k], outs["1"][k]
        assert np.isfinite(b).all()
        assert np.linalg.norm(a - b) <= 2e-3 * np.linalg.norm(a), k
