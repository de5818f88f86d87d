"""clip_grad_norm_ + AdamW kernels (csrc/optim.cu, engine.py:76-80) through the C ABI against the numpy oracle and
against torch.optim.AdamW / torch.nn.utils.clip_grad_norm_ on the same GPU.  Tolerances: fp32, same operation
order; 1e-5 relative covers fma contraction and the norm's summation order."""
import numpy as np
import pytest
import torch

from oracle import optim_oracle
from sound_event_detection_transformer_b200.optim import FusedAdamW, clip_grad_norm_

pytestmark = pytest.mark.gpu

HYPER = [dict(lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4),
         dict(lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)]
SIZES = [(64, 3, 7, 7), (256,), (11, 256), (1,), (4097,), (2048, 256), (3,), (8191,)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(*s, generator=g) for s in SIZES]


def _flat_views(tensors, dev, misalign=0):
    """views into one flat buffer (the gradient bucket of the training step); misalign > 0 makes them 4-byte aligned only"""
    n = sum(t.numel() for t in tensors) + misalign
    flat = torch.empty(n, device=dev)
    out, o = [], misalign
    for t in tensors:
        v = flat[o:o + t.numel()].view(t.shape)
        v.copy_(t)
        out.append(v)
        o += t.numel()
    return out


@pytest.mark.parametrize("misalign", [0, 1])
def test_fused_adamw_vs_oracle_and_torch(misalign):
    dev = torch.device("cuda")
    init = _params(0)
    group_of = [1, 0, 0, 1, 0, 0, 1, 0]
    mine = [torch.nn.Parameter(t.to(dev)) for t in init]
    ref = [torch.nn.Parameter(t.to(dev)) for t in init]
    fused = FusedAdamW([{"params": [p for p, g in zip(mine, group_of) if g == j], **HYPER[j]} for j in (0, 1)])
    stock = torch.optim.AdamW([{"params": [p for p, g in zip(ref, group_of) if g == j], **HYPER[j]} for j in (0, 1)])
    np_p = [t.numpy().copy() for t in init]
    state = [dict(step=0, exp_avg=np.zeros_like(a), exp_avg_sq=np.zeros_like(a)) for a in np_p]
    for it in range(4):
        grads = [g * (10.0 if it % 2 == 0 else 1e-3) for g in _params(50 + it)]
        max_norm = 0.1 if it < 3 else 0.0
        for p, v in zip(mine, _flat_views(grads, dev, misalign)):
            p.grad = v
        for p, g in zip(ref, grads):
            p.grad = g.to(dev)
        v0 = mine[0]._version
        fused.step(max_norm=max_norm)
        assert mine[0]._version > v0                       # the runtime's weight snapshot keys on this
        if max_norm > 0:
            total = torch.nn.utils.clip_grad_norm_(ref, max_norm)
            assert abs(float(fused.last_grad_norm) - float(total)) <= 1e-5 * float(total)
        stock.step()
        want_norm = optim_oracle.adamw_step(np_p, [g.numpy() for g in grads], state, group_of, HYPER, max_norm=max_norm)
        if max_norm > 0:
            assert abs(float(fused.last_grad_norm) - float(want_norm)) <= 1e-5 * float(want_norm)
        for a, p, r in zip(np_p, mine, ref):
            got = p.detach().cpu().numpy()
            np.testing.assert_allclose(got, a, rtol=1e-5, atol=1e-8)
            np.testing.assert_allclose(got, r.detach().cpu().numpy(), rtol=1e-5, atol=1e-8)
    assert fused.table_builds == 4                          # fresh gradient buffers each step here
    for st, p in zip(state, mine):
        np.testing.assert_allclose(fused.state[p]["exp_avg"].cpu().numpy(), st["exp_avg"], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(fused.state[p]["exp_avg_sq"].cpu().numpy(), st["exp_avg_sq"], rtol=1e-4, atol=1e-12)
        assert float(fused.state[p]["step"]) == 4.0


def test_state_dict_round_trip_with_torch_adamw():
    dev = torch.device("cuda")
    ps = [torch.nn.Parameter(t.to(dev)) for t in _params(3)[:3]]
    fused = FusedAdamW(ps, lr=1e-4, weight_decay=1e-4)
    for p, g in zip(ps, _params(4)):
        p.grad = g.to(dev)
    fused.step(max_norm=0.1)
    stock = torch.optim.AdamW(ps, lr=1e-4, weight_decay=1e-4)
    stock.load_state_dict(fused.state_dict())              # same keys: step / exp_avg / exp_avg_sq
    assert float(stock.state[ps[0]]["step"]) == 1.0
    torch.testing.assert_close(stock.state[ps[1]]["exp_avg"], fused.state[ps[1]]["exp_avg"])


def test_clip_grad_norm_matches_torch():
    dev = torch.device("cuda")
    a = [torch.nn.Parameter(t.to(dev)) for t in _params(5)]
    b = [torch.nn.Parameter(t.to(dev)) for t in _params(5)]
    for scale in (10.0, 1e-4):
        for p, q, g in zip(a, b, _params(6)):
            p.grad = (g * scale).to(dev)
            q.grad = (g * scale).to(dev)
        got = clip_grad_norm_(a, 0.1)
        want = torch.nn.utils.clip_grad_norm_(b, 0.1)
        assert abs(float(got) - float(want)) <= 1e-5 * float(want)
        for p, q in zip(a, b):
            torch.testing.assert_close(p.grad, q.grad, rtol=1e-5, atol=1e-10)


def test_cpu_tensors_raise():
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        FusedAdamW([p]).step()


def test_fused_adamw_mixed_step_counts_and_nan_norm():
    """A parameter whose gradient first appears later keeps its own bias correction (torch.optim.AdamW: per-parameter
    `step`); a NaN gradient norm propagates to every parameter like torch's clip_grad_norm_."""
    from sound_event_detection_transformer_b200.optim import FusedAdamW
    torch.manual_seed(0)
    a = [torch.nn.Parameter(torch.randn(300, 7, device="cuda")), torch.nn.Parameter(torch.randn(64, device="cuda"))]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    fused = FusedAdamW(a, lr=1e-2, weight_decay=1e-2)
    ref = torch.optim.AdamW(b, lr=1e-2, weight_decay=1e-2)
    for step in range(4):
        for i, (p, q) in enumerate(zip(a, b)):
            if i == 1 and step < 2:                 # the second parameter gets its first gradient at step 2
                p.grad = q.grad = None
                continue
            g = torch.randn_like(p)
            p.grad, q.grad = g.clone(), g.clone()
        torch.nn.utils.clip_grad_norm_(b, 0.1)
        ref.step()
        fused.step(max_norm=0.1)
    torch.cuda.synchronize()
    for p, q in zip(a, b):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6)
    assert float(fused.state[a[0]]["step"]) == 4.0 and float(fused.state[a[1]]["step"]) == 2.0
    a[0].grad = torch.full_like(a[0], float("nan"))
    a[1].grad = torch.zeros_like(a[1])
    fused.step(max_norm=0.1)
    torch.cuda.synchronize()
    assert torch.isnan(a[0]).all() and torch.isnan(a[1]).all()
