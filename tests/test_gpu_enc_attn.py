"""Fused encoder self-attention block (csrc/enc_attn_fused.cu) against a torch restatement of
nn.MultiheadAttention's eager path (torch/nn/functional.py:5833-5867, :6630-6659) + the residual add of
sedt/transformer.py:192-198, with bf16 rounding at the kernel's storage points (Q, K, V, the un-normalised softmax
numerators, the attention output).  Called through the C ABI (sedt_op_enc_attn).  GPU only."""
import math

import pytest
import torch

from sound_event_detection_transformer_b200 import _lib

pytestmark = pytest.mark.gpu


def r16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def reference(na, nap, w_in, b_in, w_out, b_out, kpm, x, B, S):
    d, H, hd = 256, 8, 32
    na, nap = na.float().view(B, S, d), nap.float().view(B, S, d)
    wq, wk, wv = r16(w_in).chunk(3)
    bq, bk, bv = b_in.chunk(3)
    q = r16(nap @ wq.T + bq).view(B, S, H, hd).transpose(1, 2)
    k = r16(nap @ wk.T + bk).view(B, S, H, hd).transpose(1, 2)
    v = r16(na @ wv.T + bv).view(B, S, H, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * math.sqrt(1.0 / hd)
    if kpm is not None:
        s = s.masked_fill(kpm.bool().view(B, 1, 1, S), float("-inf"))
    p = torch.exp(s - s.max(-1, keepdim=True).values)
    o = r16((r16(p) @ v) / p.sum(-1, keepdim=True))
    o = o.transpose(1, 2).reshape(B * S, d)
    return x + o @ r16(w_out).T + b_out


def run(na, nap, w_in, b_in, w_out, b_out, kpm, x, B, S, ln=None):
    """ln = (gamma, beta): also returns LayerNorm(updated x) as the kernel's bf16 side output (the layer's norm2)."""
    lib = _lib.load()
    y = x.clone()
    ln_out = torch.full((B * S, 256), float("nan"), dtype=torch.bfloat16, device="cuda") if ln is not None else None
    _lib.check(lib.sedt_op_enc_attn(na.data_ptr(), nap.data_ptr(), w_in.data_ptr(), b_in.data_ptr(), w_out.data_ptr(),
                                    b_out.data_ptr(), _lib.ptr(kpm) or None, y.data_ptr(), B, S,
                                    ln[0].data_ptr() if ln is not None else None, ln[1].data_ptr() if ln is not None else None,
                                    _lib.ptr(ln_out) or None, _lib.current_stream()))
    torch.cuda.synchronize()
    return y if ln is None else (y, ln_out)


def make(B, S, seed, masked):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g, device="cuda") * std
    na = rn(B * S, 256).to(torch.bfloat16)
    nap = (na.float() + rn(B * S, 256, std=0.7)).to(torch.bfloat16)
    w_in = rn(768, 256, std=0.08).to(torch.bfloat16)
    w_out = rn(256, 256, std=0.06).to(torch.bfloat16)
    b_in, b_out = rn(768, std=0.1), rn(256, std=0.1)
    x = rn(B * S, 256)
    kpm = None
    if masked:                                   # ragged clips: the last n_b keys of clip b are padding (at least one valid key)
        kpm = torch.zeros(B, S, dtype=torch.uint8, device="cuda")
        for b in range(B):
            n = int(torch.randint(0, max(1, S - 1), (1,), generator=g, device="cuda"))
            if n:
                kpm[b, S - n:] = 1
    return na, nap, w_in, b_in, w_out, b_out, kpm, x


@pytest.mark.parametrize("B,S,masked", [(3, 124, False), (2, 128, False), (5, 128, True), (4, 21, False), (3, 11, False),
                                         (6, 77, True), (1, 1, False), (300, 124, False), (333, 100, True)])
def test_enc_attn_fused_matches_torch(B, S, masked):
    args = make(B, S, 100 + B + S, masked)
    got = run(*args, B, S)
    want = reference(*args, B, S)
    assert torch.isfinite(got).all()
    err = ((got - want).norm() / want.norm()).item()
    worst = (got - want).abs().max().item()
    # the attention update alone (without the residual that dominates the norm)
    upd = ((got - args[-1]) - (want - args[-1])).norm() / (want - args[-1]).norm()
    print(f"enc_attn_fused B={B} S={S} masked={masked}: rel-L2 {err:.2e} (update alone {upd.item():.2e}), max abs {worst:.2e}")
    assert err < 1e-3 and upd.item() < 4e-3 and worst < 2e-2


@pytest.mark.parametrize("B,S,masked", [(3, 124, False), (5, 128, True), (200, 77, False), (2, 64, True)])
def test_enc_attn_fused_layernorm_side_output(B, S, masked):
    """The optional norm2 output: LayerNorm (eps 1e-5, two-pass statistics like layernorm_kernel) of the rows the same launch
    wrote, rounded to bf16; the main output must not change."""
    args = make(B, S, 300 + B + S, masked)
    g = torch.Generator(device="cuda").manual_seed(B * S)
    gamma = 1.0 + 0.2 * torch.randn(256, generator=g, device="cuda")
    beta = 0.1 * torch.randn(256, generator=g, device="cuda")
    y0 = run(*args, B, S)
    y, ln = run(*args, B, S, ln=(gamma, beta))
    assert torch.equal(y, y0)
    want = r16(torch.nn.functional.layer_norm(y, (256,), gamma, beta, 1e-5))
    got = ln.float()
    assert torch.isfinite(got).all()
    err = ((got - want).norm() / want.norm()).item()
    print(f"enc_attn_fused LN side output B={B} S={S}: rel-L2 {err:.2e}, max abs {(got - want).abs().max().item():.2e}")
    assert err < 2e-3 and (got - want).abs().max().item() < 6e-2          # bf16 rounding flips of 1 ulp at |y| ~ 4


def test_enc_attn_fused_is_deterministic_and_touches_only_its_rows():
    B, S = 150, 124                               # 150 clips on 148 SMs: two CTAs take a second clip
    args = make(B, S, 7, False)
    a = run(*args, B, S)
    b = run(*args, B, S)
    assert torch.equal(a, b)
    # a guard band behind the tensor must stay untouched (the last tile's rows 124..127 are out of bounds)
    x = args[-1]
    big = torch.full((B * S + 64, 256), 7.0, device="cuda")
    big[:B * S] = x
    lib = _lib.load()
    na, nap, w_in, b_in, w_out, b_out, kpm, _ = args
    na_p, nap_p = (torch.cat([t, torch.zeros(64, 256, dtype=t.dtype, device="cuda")]) for t in (na, nap))
    _lib.check(lib.sedt_op_enc_attn(na_p.data_ptr(), nap_p.data_ptr(), w_in.data_ptr(), b_in.data_ptr(), w_out.data_ptr(),
                                    b_out.data_ptr(), None, big.data_ptr(), B, S, None, None, None, _lib.current_stream()))
    torch.cuda.synchronize()
    assert torch.equal(big[:B * S], a) and bool((big[B * S:] == 7.0).all())
