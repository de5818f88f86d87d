"""Pins oracle/optim_oracle.py (numpy restatement of clip_grad_norm_ + AdamW, engine.py:76-80) against torch's own
CPU implementation -- the third-party code the reference calls."""
import numpy as np
import torch

from oracle import optim_oracle


def _case(seed, sizes=((7, 5), (33,), (4, 3, 3, 3), (1,), (4099,))):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(*s, generator=g) for s in sizes]


def test_oracle_matches_torch_adamw_and_clip():
    params = [torch.nn.Parameter(t.clone()) for t in _case(0)]
    group_of = [0, 0, 1, 1, 0]
    hyper = [dict(lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4),
             dict(lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)]
    opt = torch.optim.AdamW([{"params": [p for p, g in zip(params, group_of) if g == j], **hyper[j]} for j in (0, 1)],
                            foreach=False)
    np_p = [p.detach().numpy().copy() for p in params]
    state = [dict(step=0, exp_avg=np.zeros_like(a), exp_avg_sq=np.zeros_like(a)) for a in np_p]
    for it in range(5):
        grads = _case(100 + it)
        scale = 10.0 if it % 2 == 0 else 1e-3           # clipped and unclipped steps
        for p, g in zip(params, grads):
            p.grad = (g * scale).clone()
        np_g = [p.grad.numpy().copy() for p in params]
        total = torch.nn.utils.clip_grad_norm_(params, 0.1)
        opt.step()
        got = optim_oracle.adamw_step(np_p, np_g, state, group_of, hyper, max_norm=0.1)
        assert abs(float(got) - float(total)) <= 1e-6 * float(total)
        for a, p in zip(np_p, params):
            np.testing.assert_allclose(a, p.detach().numpy(), rtol=2e-6, atol=1e-9)
    for st, p in zip(state, params):
        np.testing.assert_allclose(st["exp_avg"], opt.state[p]["exp_avg"].numpy(), rtol=2e-5, atol=1e-9)
        np.testing.assert_allclose(st["exp_avg_sq"], opt.state[p]["exp_avg_sq"].numpy(), rtol=2e-5, atol=1e-12)


def test_oracle_no_clip_is_plain_adamw():
    params = [torch.nn.Parameter(t.clone()) for t in _case(1)]
    opt = torch.optim.AdamW(params, lr=3e-4, weight_decay=1e-2, foreach=False)
    np_p = [p.detach().numpy().copy() for p in params]
    state = [dict(step=0, exp_avg=np.zeros_like(a), exp_avg_sq=np.zeros_like(a)) for a in np_p]
    for it in range(3):
        grads = _case(7 + it)
        for p, g in zip(params, grads):
            p.grad = g.clone()
        opt.step()
        optim_oracle.adamw_step(np_p, [g.numpy() for g in grads], state, [0] * len(params),
                                [dict(lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)])
    for a, p in zip(np_p, params):
        np.testing.assert_allclose(a, p.detach().numpy(), rtol=2e-6, atol=1e-9)
