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
