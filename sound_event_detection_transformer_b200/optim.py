"""clip_grad_norm_ + AdamW of the reference's training loop (engine.py:76-80, train_sedt.py:234-240,269-270) as
two kernel launches (csrc/optim.cu) instead of ~10 foreach passes over ~200 tensors.

    optimizer = FusedAdamW(param_dicts, lr=args.lr, weight_decay=args.weight_decay)     # same arguments as AdamW
    ...
    losses.backward()
    optimizer.step(max_norm=0.1)        # == clip_grad_norm_(model.parameters(), 0.1); optimizer.step()

`clip_grad_norm_(parameters, max_norm)` is also provided with torch.nn.utils' contract (scales .grad in place and
returns the total norm as a device tensor) for callers that keep the two calls separate.  State (`step`, `exp_avg`,
`exp_avg_sq`) uses torch.optim.AdamW's keys, so optimizer state_dicts are interchangeable with the reference's
checkpoints (train_sedt.py:319).  There is no CPU path: parameters must be fp32 CUDA tensors."""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, List, Optional

import torch

from . import _lib


class _Table:
    """Device-side (tensor, chunk) table for one fixed list of (param, grad, exp_avg, exp_avg_sq, group)."""

    def __init__(self, entries, dev):
        lib = _lib.load()
        chunk = lib.sedt_optim_chunk_elems()
        arr = (_lib.SedtOptimTensor * max(len(entries), 1))()
        chunks: List[int] = []
        for i, (p, g, m, v, grp) in enumerate(entries):
            arr[i].param, arr[i].grad = p.data_ptr(), g.data_ptr()
            arr[i].exp_avg = 0 if m is None else m.data_ptr()
            arr[i].exp_avg_sq = 0 if v is None else v.data_ptr()
            arr[i].numel, arr[i].group = p.numel(), grp
            for c in range((p.numel() + chunk - 1) // chunk):
                chunks += [i, c]
        self.n = len(chunks) // 2
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.tensors = raw.to(dev)
        self.chunks = torch.tensor(chunks if chunks else [0, 0], dtype=torch.int32).to(dev)
        self.partials = torch.empty(max(self.n, 1), dtype=torch.float32, device=dev)
        self.norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.key = _table_key(entries)


def _table_key(entries):
    """Addresses AND sizes AND group ids: the table stores numel / group at build time, and the allocator may hand a freed
    address to a tensor of another size (a second model in the same process)."""
    return tuple((p.data_ptr(), g.data_ptr(), p.numel(), grp) for p, g, _, _, grp in entries)


def _check(p: torch.Tensor, g: torch.Tensor):
    if not (p.is_cuda and g.is_cuda):
        raise RuntimeError("FusedAdamW / clip_grad_norm_ need CUDA tensors (there is no CPU path)")
    if p.dtype != torch.float32 or g.dtype != torch.float32 or not p.is_contiguous() or not g.is_contiguous():
        raise RuntimeError("FusedAdamW / clip_grad_norm_ expect contiguous fp32 parameters and gradients")


_clip_tables = {}


@torch.no_grad()
def clip_grad_norm_(parameters: Iterable[torch.Tensor], max_norm: float) -> torch.Tensor:
    """torch.nn.utils.clip_grad_norm_(parameters, max_norm) (L2, engine.py:76-77) without a host synchronisation."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    ps = [p for p in parameters if p.grad is not None]
    if not ps:
        return torch.zeros([])
    for p in ps:
        _check(p, p.grad)
    dev = ps[0].device
    ent = [(p, p.grad, None, None, 0) for p in ps]
    key = _table_key(ent)
    tab = _clip_tables.get(key)
    if tab is None:
        _clip_tables.clear()                      # one live table: the training loop clips the same list every step
        tab = _clip_tables[key] = _Table(ent, dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        s = _lib.current_stream()
        _lib.check(lib.sedt_grad_norm(tab.tensors.data_ptr(), tab.chunks.data_ptr(), tab.n, tab.partials.data_ptr(),
                                      tab.norm.data_ptr(), s))
        _lib.check(lib.sedt_clip_grads(tab.tensors.data_ptr(), tab.chunks.data_ptr(), tab.n, tab.norm.data_ptr(),
                                       float(max_norm), s))
    return tab.norm[0].clone()


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW(params, lr, betas, eps, weight_decay) (amsgrad / maximize unsupported) whose step() is one
    gradient-norm pass plus one fused update pass over all parameter groups."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if not 0.5 < betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"betas {betas}: need 0.5 < beta1 < 1 and 0 <= beta2 < 1")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) > 8:
            raise ValueError("FusedAdamW supports up to 8 parameter groups")
        self._table: Optional[_Table] = None
        self.table_builds = 0            # stays at 1 when .grad keeps its addresses (flat gradient bucket)
        self.last_grad_norm: Optional[torch.Tensor] = None       # device scalar of the last step(max_norm > 0)

    def _entries(self):
        """(param, grad, exp_avg, exp_avg_sq, virtual group) per parameter with a gradient, plus the hyper-parameters of
        every virtual group.  torch.optim.AdamW keeps `step` per parameter: parameters of one param_group whose step counts
        differ (a gradient that first appears mid-run, e.g. weak_class_embed once the weak loss is switched on, or a
        checkpoint with mixed steps) get their own virtual group, i.e. their own bias correction."""
        ent, vgroups, steps = [], {}, []
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check(p, p.grad)
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                step = float(st["step"])
                vg = vgroups.setdefault((gi, step), len(vgroups))
                ent.append((p, p.grad, st["exp_avg"], st["exp_avg_sq"], vg))
                steps.append(st["step"])
        if len(vgroups) > 8:
            raise ValueError(f"FusedAdamW: {len(vgroups)} distinct (param_group, step) pairs; the kernel takes up to 8")
        return ent, vgroups, steps

    @torch.no_grad()
    def step(self, closure=None, max_norm: float = 0.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        ent, vgroups, steps = self._entries()
        if not ent:
            return loss
        dev = ent[0][0].device
        key = _table_key(ent)
        if self._table is None or self._table.key != key or \
                self._table_state != tuple((m.data_ptr(), v.data_ptr()) for _, _, m, v, _ in ent):
            self._table = _Table(ent, dev)
            self.table_builds += 1
            self._table_state = tuple((m.data_ptr(), v.data_ptr()) for _, _, m, v, _ in ent)
        tab = self._table
        groups = (_lib.SedtAdamWGroup * len(vgroups))()
        for (gi, step0), vg in vgroups.items():
            group = self.param_groups[gi]
            step = step0 + 1.0
            b1, b2 = group["betas"]
            lr = float(group["lr"])
            groups[vg].decay = 1.0 - lr * group["weight_decay"]
            groups[vg].w1, groups[vg].beta2, groups[vg].w2 = 1.0 - b1, b2, 1.0 - b2
            groups[vg].bc2_sqrt = math.sqrt(1.0 - b2 ** step)
            groups[vg].eps = group["eps"]
            groups[vg].neg_step = -(lr / (1.0 - b1 ** step))
        torch._foreach_add_(steps, 1.0)
        lib = _lib.load()
        with torch.cuda.device(dev):
            s = _lib.current_stream()
            norm_ptr = None
            if max_norm and max_norm > 0:
                _lib.check(lib.sedt_grad_norm(tab.tensors.data_ptr(), tab.chunks.data_ptr(), tab.n,
                                              tab.partials.data_ptr(), tab.norm.data_ptr(), s))
                norm_ptr = tab.norm.data_ptr()
                self.last_grad_norm = tab.norm
            _lib.check(lib.sedt_adamw_step(tab.tensors.data_ptr(), tab.chunks.data_ptr(), tab.n, groups,
                                           len(vgroups), norm_ptr, float(max_norm or 0.0), s))
        # the update wrote the parameters behind autograd's back: bump the version counters so that the runtime
        # re-packs its bf16 weight snapshot (runtime.ensure_packed keys on _version)
        self._bump(ent)
        return loss

    @staticmethod
    def _bump(ent):
        ps = [p for p, *_ in ent]
        try:        # no kernel launch: just the counters
            torch._C._autograd._unsafe_set_version_counter(ps, [p._version + 1 for p in ps])
        except Exception:
            torch._foreach_add_(ps, 0.0)
