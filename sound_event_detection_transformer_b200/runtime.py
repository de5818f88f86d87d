"""Python owner of one native model handle: collects the module's parameters
into the library's weight table, keeps the packed-weight snapshot fresh,
caches the workspace and drives sedt_forward on the current CUDA stream.
PyTorch is used for device memory and streams only."""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, Optional

import torch

from . import _lib


class _TrainSlot:
    """Resources of one forward_train / backward pair in flight (see ForwardRuntime._acquire_slot)."""

    def __init__(self, index: int, seed: int):
        self.index, self.seed = index, seed
        self.tape: Optional[torch.Tensor] = None
        self.owner = None          # weakref to the token of the forward that holds the slot
        self.gen = 0               # bumped by every forward_train on this slot
        self.g: Optional[dict] = None      # graph mode: static buffers + captured graphs

    def busy(self) -> bool:
        return self.owner is not None and self.owner() is not None


class _TrainCtx:
    __slots__ = ("slot", "gen", "x", "m8", "B", "T", "F", "dropout", "graph", "P", "PT")

    def __init__(self, slot, gen, x, m8, B, T, F, dropout, graph, P=0, PT=0):
        self.slot, self.gen, self.x, self.m8, self.B, self.T, self.F, self.dropout, self.graph = slot, gen, x, m8, B, T, F, dropout, graph
        self.P, self.PT = P, PT


class ForwardRuntime:
    def __init__(self, cfg: Dict[str, int]):
        self.lib = _lib.load()
        self.cfg = _lib.SedtConfig(**cfg)
        self.handle = C.c_void_p()
        _lib.check(self.lib.sedt_model_create(C.byref(self.cfg), C.byref(self.handle)))
        n = self.lib.sedt_model_num_weights(self.handle)
        self.names = [self.lib.sedt_model_weight_name(self.handle, i).decode() for i in range(n)]
        self.numels = [self.lib.sedt_model_weight_numel(self.handle, i) for i in range(n)]
        self.packed_bytes = self.lib.sedt_model_packed_bytes(self.handle)
        self.packed: Optional[torch.Tensor] = None
        self._stamp = None
        self._ws: Optional[torch.Tensor] = None
        self._keep = None
        # CUDA-graph replay of the whole forward (one graph per input shape); see forward(use_graph=True)
        self._graphs: Dict[tuple, dict] = {}
        self.graph_kernel_launches = 0     # kernels executed through graph replays (the library counter only sees eager launches)
        self.max_train_slots = 4           # forward_train calls that may wait for their backward at the same time (per input shape)

    def __del__(self):
        try:
            if self.handle:
                self.lib.sedt_model_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------
    def ensure_packed(self, tensors: Dict[str, torch.Tensor], use_graph: bool = False) -> None:
        """tensors: reference state_dict name -> live parameter/buffer (CUDA, fp32).  use_graph (training loop: the
        optimizer rewrites the same tensors in place every step) replays the ~400 small packing kernels as one
        CUDA graph keyed by the parameter addresses."""
        stamp = tuple((tensors[n].data_ptr(), tensors[n]._version) for n in self.names)
        if stamp == self._stamp and self.packed is not None:
            return
        pkey = tuple(p for p, _ in stamp)
        if use_graph and self.packed is not None and getattr(self, "_pack_graph_key", None) == pkey:
            self._pack_graph.replay()
            self._stamp = stamp
            return
        dev = None
        ptrs = (C.c_void_p * len(self.names))()
        keep = []
        for i, n in enumerate(self.names):
            t = tensors[n]
            if not t.is_cuda:
                raise RuntimeError(f"parameter {n} lives on {t.device}: move the model to a CUDA device "
                                   "(there is no CPU path)")
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.detach().to(torch.float32).contiguous()
                keep.append(t)
            if t.numel() != self.numels[i]:
                raise RuntimeError(f"parameter {n} has {t.numel()} elements, expected {self.numels[i]}")
            dev = t.device
            ptrs[i] = t.data_ptr()
        realloc = self.packed is None or self.packed.device != dev
        if realloc:
            self.packed = torch.empty(self.packed_bytes + 256, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self.lib.sedt_model_pack(self.handle, ptrs, self._aligned(self.packed), self.packed_bytes,
                                                _lib.current_stream()))
        self._keep = keep
        self._ptrs = ptrs
        self._stamp = stamp
        # graphs bake in the packed-buffer and parameter addresses (not the values): re-capture only when those move
        if realloc or keep or getattr(self, "_ptr_key", None) != pkey:
            self._graphs.clear()
            self._pack_graph_key = None
            self._pack_epoch = getattr(self, "_pack_epoch", 0) + 1      # train slots re-capture when this moves
        self._ptr_key = pkey
        if use_graph and not keep and self._pack_graph_key != pkey:
            with torch.cuda.device(dev):
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    _lib.check(self.lib.sedt_model_pack(self.handle, ptrs, self._aligned(self.packed), self.packed_bytes,
                                                        _lib.current_stream()))
            self._pack_graph, self._pack_graph_key = g, pkey

    @staticmethod
    def _aligned(buf: torch.Tensor) -> int:
        return (buf.data_ptr() + 255) & ~255

    # ---- forward -------------------------------------------------------------
    def feature_shape(self, T: int, F: int):
        h, w = C.c_int(), C.c_int()
        _lib.check(self.lib.sedt_feature_shape(T, F, self.cfg.dilation, C.byref(h), C.byref(w)))
        return h.value, w.value

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor], patches: Optional[torch.Tensor] = None,
                want_memory: bool = False, want_feat: bool = False, use_graph: bool = False) -> Dict[str, torch.Tensor]:
        """One forward.  use_graph=True replays a CUDA graph of the launch sequence captured for this input
        shape (static input / output buffers owned by the runtime; the returned tensors are only valid until
        the next call with the same shape)."""
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 1
        if use_graph and not want_memory and not want_feat:
            return self._forward_graph(x, mask, patches)
        return self._forward_eager(x, mask, patches, want_memory, want_feat)

    def _alloc_outputs(self, B, T, F, P, dev, want_memory=False, want_feat=False):
        cfg = self.cfg
        D = cfg.dec_layers
        if cfg.self_sup:
            qall = q = P * (cfg.num_queries // cfg.num_patches)
        else:
            q = cfg.num_queries
            qall = q + (1 if cfg.dec_at else 0)
        ncls = 1 if cfg.self_sup else cfg.num_classes
        f32 = dict(dtype=torch.float32, device=dev)
        res = {
            "hs": torch.empty(D, B, qall, cfg.hidden_dim, **f32),
            "logits": torch.empty(D, B, q, ncls + 1, **f32),
            "boxes": torch.empty(D, B, q, 2, **f32),
        }
        if cfg.dec_at:
            res["at"] = torch.empty(B, ncls, **f32)
        H, W = self.feature_shape(T, F)
        if want_memory:
            res["memory"] = torch.empty(B, H * W, cfg.hidden_dim, **f32)
        if want_feat:
            res["feat"] = torch.empty(B, H, W, 2048, **f32)
        if cfg.self_sup:
            res["gt_feature"] = torch.empty(B * P, 2048, **f32)
            if cfg.feature_recon:
                res["pred_feature"] = torch.empty(D, B, q, 2048, **f32)
        return res

    def _launch(self, x, m8, patches, P, PT, ws, res):
        B, _, T, F = x.shape
        outs = _lib.SedtOutputs(**{k: _lib.ptr(res.get(k)) or None for k, _ in _lib.SedtOutputs._fields_})
        _lib.check(self.lib.sedt_forward(self.handle, x.data_ptr(), _lib.ptr(m8) or None, B, T, F,
                                         _lib.ptr(patches) or None, P, PT, self._aligned(ws), ws.numel() - 256,
                                         C.byref(outs), _lib.current_stream()))

    def _forward_eager(self, x, mask, patches, want_memory, want_feat):
        x = x.contiguous()
        B, _, T, F = x.shape
        dev = x.device
        P = PT = 0
        if self.cfg.self_sup:
            assert patches is not None and patches.dim() == 5
            patches = patches.to(dev, torch.float32).contiguous()
            P, PT = patches.shape[1], patches.shape[3]
        res = self._alloc_outputs(B, T, F, P, dev, want_memory, want_feat)
        need = self.lib.sedt_workspace_bytes(self.handle, B, T, F, P, PT)
        if need < 0:
            _lib.check(int(need))
        if self._ws is None or self._ws.numel() < need + 256 or self._ws.device != dev:
            self._ws = None
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
        m8 = None
        if mask is not None:
            m8 = mask.to(dev).contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(dev, torch.uint8).contiguous()
        with torch.cuda.device(dev):
            self._launch(x, m8, patches, P, PT, self._ws, res)
        return res

    def _forward_graph(self, x, mask, patches):
        B, _, T, F = x.shape
        dev = x.device
        P = PT = 0
        if self.cfg.self_sup:
            assert patches is not None and patches.dim() == 5
            P, PT = patches.shape[1], patches.shape[3]
        key = (B, T, F, P, PT, mask is not None, dev.index)
        g = self._graphs.get(key)
        if g is None:
            need = self.lib.sedt_workspace_bytes(self.handle, B, T, F, P, PT)
            if need < 0:
                _lib.check(int(need))
            g = {"x": torch.empty(B, 1, T, F, dtype=torch.float32, device=dev),
                 "mask": torch.empty(B, T, F, dtype=torch.uint8, device=dev) if mask is not None else None,
                 "patches": torch.empty(B, P, 1, PT, F, dtype=torch.float32, device=dev) if P else None,
                 "ws": torch.empty(need + 256, dtype=torch.uint8, device=dev),
                 "res": self._alloc_outputs(B, T, F, P, dev)}
            with torch.cuda.device(dev):
                # one eager pass first: lazy one-time setup (driver entry point, kernel attributes) must not be captured
                self._launch(g["x"], g["mask"], g["patches"], P, PT, g["ws"], g["res"])
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                n0 = self.lib.sedt_launch_count()
                with torch.cuda.graph(graph):
                    self._launch(g["x"], g["mask"], g["patches"], P, PT, g["ws"], g["res"])
                g["nlaunch"] = int(self.lib.sedt_launch_count() - n0)
            g["graph"] = graph
            self._graphs[key] = g
        g["x"].copy_(x, non_blocking=True)
        if mask is not None:
            g["mask"].copy_(mask.view(torch.uint8) if mask.dtype == torch.bool else mask, non_blocking=True)
        if P:
            g["patches"].copy_(patches, non_blocking=True)
        g["graph"].replay()
        self.graph_kernel_launches += g["nlaunch"]
        return g["res"]

    # ---- training step ---------------------------------------------------------
    def grad_layout(self):
        """(total fp32 elements of the flat gradient buffer, {state_dict name: element offset})."""
        if getattr(self, "_grad_layout", None) is None:
            n = int(self.lib.sedt_grad_numel(self.handle))
            offs = {name: int(self.lib.sedt_grad_offset(self.handle, i)) for i, name in enumerate(self.names)}
            self._grad_layout = (n, offs)
        return self._grad_layout

    def backbone_grad_range(self):
        """[lo, hi) element range of the `backbone.*` gradients in the flat buffer (they are contiguous in state_dict order)."""
        n, offs = self.grad_layout()
        bb = [i for i, name in enumerate(self.names) if name.startswith("backbone.")]
        lo = offs[self.names[bb[0]]]
        hi = offs[self.names[bb[-1] + 1]] if bb[-1] + 1 < len(self.names) else n
        return lo, hi

    def bucket_event(self, enable: bool = True):
        """The event sedt_backward records once the non-backbone gradients are final (overlapped all-reduce, see
        sedt/model.py: _allreduce_buckets); enable=False detaches it."""
        if not enable:
            _lib.check(self.lib.sedt_model_set_bucket_event(self.handle, None))
            return None
        ev = getattr(self, "_bucket_event", None)
        if ev is None:
            ev = self._bucket_event = torch.cuda.Event()
            ev.record()                                   # forces the lazy creation of the cudaEvent_t (outside any capture)
        _lib.check(self.lib.sedt_model_set_bucket_event(self.handle, ev.cuda_event))
        return ev

    # Train slots.  One forward_train / backward pair owns one slot: its own activation tape and dropout RNG stream and, in
    # graph mode, its own static input / output / gradient buffers and captured graphs.  A forward issued while earlier
    # forwards still wait for their backward (the semi-supervised loop runs the labelled and the unlabelled batch through
    # the model before ONE total_losses.backward(), engine.py:134-170) takes a fresh slot instead of overwriting the tape.
    # A slot is busy while the autograd node that owns it (its token) is alive and has not run backward.
    def _slot_seed(self, index: int) -> int:
        if getattr(self, "_seed", None) is None:
            self._seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF        # fixed per runtime: graphs bake it in
        return (self._seed + index * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF   # slot 0 keeps the runtime's seed

    def _acquire_slot(self, key, token) -> "_TrainSlot":
        slots = self._train_slots.setdefault(key, []) if hasattr(self, "_train_slots") else None
        if slots is None:
            self._train_slots = {key: []}
            slots = self._train_slots[key]
        for sl in slots:
            if not sl.busy():
                break
        else:
            if len(slots) >= self.max_train_slots:
                raise RuntimeError(f"{len(slots)} train-mode forwards are waiting for their backward: outputs of grad-enabled "
                                   "forwards are being kept alive without ever calling backward() (each holds an activation tape); "
                                   "run such forwards under torch.no_grad(), or raise runtime.max_train_slots")
            n = sum(len(v) for v in self._train_slots.values())
            sl = _TrainSlot(n, self._slot_seed(n))
            slots.append(sl)
        sl.owner = weakref.ref(token) if token is not None else None
        sl.gen += 1
        return sl

    def train_slots_in_use(self) -> int:
        return sum(1 for v in getattr(self, "_train_slots", {}).values() for sl in v if sl.busy())

    def forward_train(self, x: torch.Tensor, mask: Optional[torch.Tensor], use_graph: bool = False, dropout: float = 0.0,
                      token=None, patches: Optional[torch.Tensor] = None, query_keep: Optional[torch.Tensor] = None):
        """Forward in train mode: same outputs as forward(); the activations stay in the slot's tape until backward().
        `token`: the object whose lifetime marks the slot busy (the autograd context); None = fire and forget (a train-mode
        forward without gradients: dropout is applied, nothing is kept).  use_graph replays the launch sequence as a CUDA
        graph per (input shape, slot): static buffers, overwritten by the slot's next step.  Returns (res, ctx)."""
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 1
        x = x.contiguous()
        B, _, T, F = x.shape
        dev = x.device
        dropout = float(dropout)
        P = PT = 0
        if self.cfg.self_sup:
            # SP-SEDT pretraining step (spsedt.py:63-69): patches [B, P, 1, PT, F] and the query-drop mask [B, Q] (1 = keep);
            # eager launches only (the graph path keeps static buffers for the supervised model)
            assert patches is not None and patches.dim() == 5 and query_keep is not None
            patches = patches.to(dev, torch.float32).contiguous()
            query_keep = query_keep.to(dev, torch.uint8).contiguous()
            P, PT = int(patches.shape[1]), int(patches.shape[3])
            use_graph = False
        if use_graph:
            return self._forward_train_graph(x, mask, dropout, token)
        sl = self._acquire_slot(("eager", dev.index), token)
        need = int(self.lib.sedt_train_tape_bytes_sp(self.handle, B, T, F, int(mask is not None), P, PT))
        if need < 0:
            _lib.check(need)
        if sl.tape is None or sl.tape.numel() < need + 256 or sl.tape.device != dev:
            sl.tape = None
            sl.tape = torch.empty(need + 256, dtype=torch.uint8, device=dev)
        m8 = None
        if mask is not None:
            m8 = mask.to(dev).contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(dev, torch.uint8).contiguous()
        res = self._alloc_outputs(B, T, F, P, dev)
        outs = _lib.SedtOutputs(**{k: _lib.ptr(res.get(k)) or None for k, _ in _lib.SedtOutputs._fields_})
        with torch.cuda.device(dev):
            if self.cfg.self_sup:
                _lib.check(self.lib.sedt_forward_train_sp(self.handle, x.data_ptr(), _lib.ptr(m8) or None, B, T, F, patches.data_ptr(),
                                                          P, PT, query_keep.data_ptr(), self._aligned(sl.tape), sl.tape.numel() - 256,
                                                          C.byref(outs), dropout, sl.seed, _lib.current_stream()))
            else:
                _lib.check(self.lib.sedt_forward_train(self.handle, x.data_ptr(), _lib.ptr(m8) or None, B, T, F, self._aligned(sl.tape),
                                                       sl.tape.numel() - 256, C.byref(outs), dropout, sl.seed,
                                                       _lib.current_stream()))
        return res, _TrainCtx(sl, sl.gen, x, m8, B, T, F, dropout, False, P, PT)

    def _forward_train_graph(self, x, mask, dropout, token):
        B, _, T, F = x.shape
        dev = x.device
        has_mask = mask is not None
        sl = self._acquire_slot((B, T, F, has_mask, dev.index, dropout), token)
        g = sl.g
        if g is None:
            need = int(self.lib.sedt_train_tape_bytes(self.handle, B, T, F, int(has_mask)))
            if need < 0:
                _lib.check(need)
            sl.tape = torch.empty(need + 256, dtype=torch.uint8, device=dev)
            g = sl.g = {"x": torch.empty(B, 1, T, F, dtype=torch.float32, device=dev),
                        "mask": torch.empty(B, T, F, dtype=torch.uint8, device=dev) if has_mask else None,
                        "res": self._alloc_outputs(B, T, F, 0, dev), "fwd": None, "bwd": None, "epoch": self._pack_epoch}
        elif g["epoch"] != self._pack_epoch:           # the parameters / packed buffer moved: the captured addresses are stale
            g["fwd"] = g["bwd"] = None
            g["epoch"] = self._pack_epoch
        g["x"].copy_(x, non_blocking=True)
        if has_mask:
            g["mask"].copy_(mask.view(torch.uint8) if mask.dtype == torch.bool else mask, non_blocking=True)
        if g["fwd"] is None:
            res, tape = g["res"], sl.tape
            outs = _lib.SedtOutputs(**{k: _lib.ptr(res.get(k)) or None for k, _ in _lib.SedtOutputs._fields_})

            def launch():
                _lib.check(self.lib.sedt_forward_train(self.handle, g["x"].data_ptr(), _lib.ptr(g["mask"]) or None, B, T, F,
                                                       self._aligned(tape), tape.numel() - 256, C.byref(outs), dropout,
                                                       sl.seed, _lib.current_stream()))
            with torch.cuda.device(dev):
                launch()                                  # lazy one-time setup must not be captured
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                n0 = self.lib.sedt_launch_count()
                with torch.cuda.graph(graph):
                    launch()
                g["fwd_launches"] = int(self.lib.sedt_launch_count() - n0)
            g["fwd"] = graph
        g["fwd"].replay()
        self.graph_kernel_launches += g["fwd_launches"]
        return g["res"], _TrainCtx(sl, sl.gen, g["x"], g["mask"], B, T, F, dropout, True)

    def _backward_graph(self, ctx, d_logits, d_boxes, d_at, train_backbone, post=None):
        """post(grads): optional work captured into the same graph right after the backward launches (the data-parallel
        gradient all-reduce: NCCL collectives are capturable).  If capturing it fails the graph is re-captured without it and
        g["post_captured"] stays False: the caller then runs it eagerly after the replay."""
        sl, x, m8, B, T, F = ctx.slot, ctx.x, ctx.m8, ctx.B, ctx.T, ctx.F
        g = sl.g
        dev = x.device
        if g["bwd"] is None or g.get("bwd_tb") != bool(train_backbone) or g.get("bwd_post") != (post is not None):
            need = int(self.lib.sedt_backward_workspace_bytes(self.handle, B, T, F))
            if need < 0:
                _lib.check(need)
            n, _ = self.grad_layout()
            res = g["res"]
            g["ws"] = torch.empty(need + 256, dtype=torch.uint8, device=dev)
            g["flat"] = torch.zeros(n + 64, dtype=torch.float32, device=dev)
            g["dl"], g["db"] = torch.zeros_like(res["logits"]), torch.zeros_like(res["boxes"])
            g["da"] = torch.zeros_like(res["at"]) if "at" in res else None
            shift = ((-g["flat"].data_ptr()) % 256) // 4
            g["grads"] = g["flat"][shift:shift + n]
            tape, ws = sl.tape, g["ws"]

            def launch():
                _lib.check(self.lib.sedt_backward(self.handle, self._ptrs, x.data_ptr(), _lib.ptr(m8) or None, B, T, F,
                                                  self._aligned(tape), tape.numel() - 256, self._aligned(ws), ws.numel() - 256,
                                                  g["dl"].data_ptr(), g["db"].data_ptr(), _lib.ptr(g["da"]) or None,
                                                  g["grads"].data_ptr(), int(train_backbone), ctx.dropout, _lib.current_stream()))
            with torch.cuda.device(dev):
                launch()
                if post is not None:
                    post(g["grads"])                       # warm-up outside the capture (communicator setup)
                torch.cuda.synchronize(dev)
                g["post_captured"] = False
                graph = None
                if post is not None:
                    try:
                        graph = torch.cuda.CUDAGraph()
                        n0 = self.lib.sedt_launch_count()
                        with torch.cuda.graph(graph):
                            launch()
                            post(g["grads"])
                        g["post_captured"] = True
                    except Exception as exc:               # e.g. a collective backend that cannot be captured
                        import warnings
                        warnings.warn(f"gradient all-reduce could not be captured into the backward graph ({exc!r}); "
                                      "it runs eagerly after the replay")
                        torch.cuda.synchronize(dev)
                        graph = None
                if graph is None:
                    graph = torch.cuda.CUDAGraph()
                    n0 = self.lib.sedt_launch_count()
                    with torch.cuda.graph(graph):
                        launch()
                g["bwd_launches"] = int(self.lib.sedt_launch_count() - n0)
            g["bwd"], g["bwd_tb"], g["bwd_post"] = graph, bool(train_backbone), post is not None
        for dst, src in ((g["dl"], d_logits), (g["db"], d_boxes), (g["da"], d_at)):
            if dst is not None:
                if src is None:
                    dst.zero_()
                else:
                    dst.copy_(src, non_blocking=True)
        g["bwd"].replay()
        self.graph_kernel_launches += g["bwd_launches"]
        if post is not None and not g["post_captured"]:
            post(g["grads"])
        return g["grads"]

    def backward(self, ctx: "_TrainCtx", d_logits, d_boxes, d_at, train_backbone: bool, d_pred_feature=None, post=None) -> torch.Tensor:
        """Gradients of every trainable state_dict entry in one flat fp32 tensor (see grad_layout()).  In graph mode the
        tensor is the slot's static buffer: the next backward of the same slot overwrites it (callers that hand views of
        it to autograd must copy, see sedt/model.py: _TrainStep)."""
        sl = ctx.slot
        if sl.gen != ctx.gen:
            raise RuntimeError("backward: the activation tape of this forward has been overwritten by a later forward_train "
                               "(backward called twice, or after the autograd graph had been released)")
        try:
            if ctx.graph:
                return self._backward_graph(ctx, d_logits, d_boxes, d_at, train_backbone, post)
            x, m8, B, T, F = ctx.x, ctx.m8, ctx.B, ctx.T, ctx.F
            dev = x.device
            need = int(self.lib.sedt_backward_workspace_bytes_sp(self.handle, B, T, F, ctx.P, ctx.PT))
            if need < 0:
                _lib.check(need)
            ws = getattr(self, "_bws", None)
            if ws is None or ws.numel() < need + 256 or ws.device != dev:
                self._bws = None
                self._bws = ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
            n, _ = self.grad_layout()
            flat = torch.empty(n + 64, dtype=torch.float32, device=dev)
            shift = ((-flat.data_ptr()) % 256) // 4
            grads = flat[shift:shift + n]

            def f32(t):
                return None if t is None else t.detach().to(torch.float32).contiguous()
            d_logits, d_boxes, d_at, d_pred_feature = f32(d_logits), f32(d_boxes), f32(d_at), f32(d_pred_feature)
            tape = sl.tape
            if self.cfg.self_sup:
                with torch.cuda.device(dev):
                    _lib.check(self.lib.sedt_backward_sp(self.handle, self._ptrs, x.data_ptr(), _lib.ptr(m8) or None, B, T, F, ctx.P, ctx.PT,
                                                         self._aligned(tape), tape.numel() - 256, self._aligned(ws), ws.numel() - 256,
                                                         _lib.ptr(d_logits) or None, _lib.ptr(d_boxes) or None,
                                                         _lib.ptr(d_pred_feature) or None, grads.data_ptr(), ctx.dropout,
                                                         _lib.current_stream()))
                if post is not None:
                    post(grads)
                return grads
            with torch.cuda.device(dev):
                _lib.check(self.lib.sedt_backward(self.handle, self._ptrs, x.data_ptr(), _lib.ptr(m8) or None, B, T, F,
                                                  self._aligned(tape), tape.numel() - 256, self._aligned(ws), ws.numel() - 256,
                                                  _lib.ptr(d_logits) or None, _lib.ptr(d_boxes) or None, _lib.ptr(d_at) or None,
                                                  grads.data_ptr(), int(train_backbone), ctx.dropout, _lib.current_stream()))
            if post is not None:
                post(grads)
            return grads
        finally:
            sl.owner = None                    # the pair is complete: the slot may serve the next forward

    def kernel_launches(self) -> int:
        """Kernels of this library executed so far on behalf of this process (eager + graph replays)."""
        return int(self.lib.sedt_launch_count()) + self.graph_kernel_launches
