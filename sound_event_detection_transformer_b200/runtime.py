"""Python owner of one native model handle: collects the module's parameters
into the library's weight table, keeps the packed-weight snapshot fresh,
caches the workspace and drives sedt_forward on the current CUDA stream.
PyTorch is used for device memory and streams only."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib


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
            self._train_graphs = {}
            self._pack_graph_key = None
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

    def forward_train(self, x: torch.Tensor, mask: Optional[torch.Tensor], use_graph: bool = False, dropout: float = 0.0):
        """Forward in train mode: same outputs as forward(); the activations stay in a runtime-owned tape until
        backward() (one forward/backward pair in flight per runtime).  use_graph replays the launch sequence as a
        CUDA graph per input shape (static input / output buffers, overwritten by the next step)."""
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 1
        x = x.contiguous()
        B, _, T, F = x.shape
        dev = x.device
        self._dropout = float(dropout)
        if getattr(self, "_seed", None) is None:
            self._seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF        # fixed per runtime: graphs bake it in
        if use_graph:
            return self._forward_train_graph(x, mask)
        need = int(self.lib.sedt_train_tape_bytes(self.handle, B, T, F, int(mask is not None)))
        if need < 0:
            _lib.check(need)
        tape = getattr(self, "_tape", None)
        if tape is None or tape.numel() < need + 256 or tape.device != dev:
            self._tape = None
            self._tape = tape = torch.empty(need + 256, dtype=torch.uint8, device=dev)
        m8 = None
        if mask is not None:
            m8 = mask.to(dev).contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(dev, torch.uint8).contiguous()
        res = self._alloc_outputs(B, T, F, 0, dev)
        outs = _lib.SedtOutputs(**{k: _lib.ptr(res.get(k)) or None for k, _ in _lib.SedtOutputs._fields_})
        with torch.cuda.device(dev):
            _lib.check(self.lib.sedt_forward_train(self.handle, x.data_ptr(), _lib.ptr(m8) or None, B, T, F, self._aligned(tape),
                                                   tape.numel() - 256, C.byref(outs), self._dropout, self._seed,
                                                   _lib.current_stream()))
        return res, (x, m8, B, T, F)

    def _train_graph_state(self, B, T, F, has_mask, dev):
        key = (B, T, F, has_mask, dev.index, self._dropout)
        g = getattr(self, "_train_graphs", {}).get(key)
        if g is None:
            if not hasattr(self, "_train_graphs"):
                self._train_graphs = {}
            need = int(self.lib.sedt_train_tape_bytes(self.handle, B, T, F, int(has_mask)))
            if need < 0:
                _lib.check(need)
            self._tape = torch.empty(need + 256, dtype=torch.uint8, device=dev)
            g = {"x": torch.empty(B, 1, T, F, dtype=torch.float32, device=dev),
                 "mask": torch.empty(B, T, F, dtype=torch.uint8, device=dev) if has_mask else None,
                 "res": self._alloc_outputs(B, T, F, 0, dev), "fwd": None, "bwd": None}
            self._train_graphs[key] = g
        return g

    def _forward_train_graph(self, x, mask):
        B, _, T, F = x.shape
        dev = x.device
        g = self._train_graph_state(B, T, F, mask is not None, dev)
        g["x"].copy_(x, non_blocking=True)
        if mask is not None:
            g["mask"].copy_(mask.view(torch.uint8) if mask.dtype == torch.bool else mask, non_blocking=True)
        if g["fwd"] is None:
            res, tape = g["res"], self._tape
            outs = _lib.SedtOutputs(**{k: _lib.ptr(res.get(k)) or None for k, _ in _lib.SedtOutputs._fields_})

            def launch():
                _lib.check(self.lib.sedt_forward_train(self.handle, g["x"].data_ptr(), _lib.ptr(g["mask"]) or None, B, T, F,
                                                       self._aligned(tape), tape.numel() - 256, C.byref(outs), self._dropout,
                                                       self._seed, _lib.current_stream()))
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
        return g["res"], (g["x"], g["mask"], B, T, F, g)

    def _backward_graph(self, ctx, d_logits, d_boxes, d_at, train_backbone):
        x, m8, B, T, F, g = ctx
        dev = x.device
        if g["bwd"] is None or g.get("bwd_tb") != bool(train_backbone):
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
            tape, ws = self._tape, g["ws"]

            def launch():
                _lib.check(self.lib.sedt_backward(self.handle, self._ptrs, x.data_ptr(), _lib.ptr(m8) or None, B, T, F,
                                                  self._aligned(tape), tape.numel() - 256, self._aligned(ws), ws.numel() - 256,
                                                  g["dl"].data_ptr(), g["db"].data_ptr(), _lib.ptr(g["da"]) or None,
                                                  g["grads"].data_ptr(), int(train_backbone), self._dropout, _lib.current_stream()))
            with torch.cuda.device(dev):
                launch()
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                n0 = self.lib.sedt_launch_count()
                with torch.cuda.graph(graph):
                    launch()
                g["bwd_launches"] = int(self.lib.sedt_launch_count() - n0)
            g["bwd"], g["bwd_tb"] = graph, bool(train_backbone)
        for dst, src in ((g["dl"], d_logits), (g["db"], d_boxes), (g["da"], d_at)):
            if dst is not None:
                if src is None:
                    dst.zero_()
                else:
                    dst.copy_(src, non_blocking=True)
        g["bwd"].replay()
        self.graph_kernel_launches += g["bwd_launches"]
        return g["grads"]

    def backward(self, ctx, d_logits, d_boxes, d_at, train_backbone: bool) -> torch.Tensor:
        """Gradients of every trainable state_dict entry in one flat fp32 tensor (see grad_layout())."""
        if len(ctx) == 6:
            return self._backward_graph(ctx, d_logits, d_boxes, d_at, train_backbone)
        x, m8, B, T, F = ctx
        dev = x.device
        need = int(self.lib.sedt_backward_workspace_bytes(self.handle, B, T, F))
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
        d_logits, d_boxes, d_at = f32(d_logits), f32(d_boxes), f32(d_at)
        tape = self._tape
        with torch.cuda.device(dev):
            _lib.check(self.lib.sedt_backward(self.handle, self._ptrs, x.data_ptr(), _lib.ptr(m8) or None, B, T, F,
                                              self._aligned(tape), tape.numel() - 256, self._aligned(ws), ws.numel() - 256,
                                              _lib.ptr(d_logits) or None, _lib.ptr(d_boxes) or None, _lib.ptr(d_at) or None,
                                              grads.data_ptr(), int(train_backbone), self._dropout, _lib.current_stream()))
        return grads

    def kernel_launches(self) -> int:
        """Kernels of this library executed so far on behalf of this process (eager + graph replays)."""
        return int(self.lib.sedt_launch_count()) + self.graph_kernel_launches
