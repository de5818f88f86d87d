"""SEDT / SPSEDT with the reference's constructor, state_dict and forward()
contract (sedt/sedt.py:17-131, sedt/spsedt.py:14-95); the forward itself is
one call into the native runtime (hand-written sm_100a kernels)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from ..runtime import ForwardRuntime
from ..utils import NestedTensor, nested_tensor_from_tensor_list
from .modules import MLP

PRECISIONS = {"fp32": 0, "bf16": 1}


class _Token:
    """Lifetime marker of one forward's autograd context: the runtime slot (tape) it holds stays busy while it is alive
    and backward has not run."""
    __slots__ = ("__weakref__",)


def _allreduce_buckets(rt, train_backbone: bool, overlap: bool):
    """The training step's one exchange: the mean all-reduce of the flat gradient buffer inside backward.
    overlap=False (default): one collective after the last backward kernel.  overlap=True: two buckets -- the gradients outside
    the backbone are averaged on a side stream as soon as sedt_backward signals them final (bucket event), while the backbone
    backward is still running; the backbone bucket follows on the main stream.  Measured on 2 x B200 (B = 64 / GPU): 9.51 ms per
    step overlapped vs ~9.25 ms sequential vs 8.94 ms without any exchange -- the persistent 148-CTA GEMM kernels of the backbone
    backward lose SMs to NCCL's CTAs and run a second wave, which costs more than the 0.3 ms the overlap hides; hence the default.
    Returns post(flat) for ForwardRuntime.backward, or None for a single process."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        rt.bucket_event(False)
        return None
    from ..parallel import allreduce_mean_
    if not overlap or not train_backbone:
        rt.bucket_event(False)
        return allreduce_mean_
    lo, hi = rt.backbone_grad_range()
    ev = rt.bucket_event(True)
    if getattr(rt, "_side_stream", None) is None:
        rt._side_stream = torch.cuda.Stream()
    side = rt._side_stream

    def post(flat):
        main = torch.cuda.current_stream()
        with torch.cuda.stream(side):
            side.wait_event(ev)                      # recorded by sedt_backward after input_proj's weight gradient
            allreduce_mean_(flat[:lo])
            if hi < flat.numel():
                allreduce_mean_(flat[hi:])
        allreduce_mean_(flat[lo:hi])                 # main stream: after the backbone backward
        main.wait_stream(side)
    return post


def _detach_grads_from(static, params, names, shapes, rt):
    if static is None:
        return
    lo, hi = static.data_ptr(), static.data_ptr() + static.numel() * 4
    if not any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in params):
        return
    keep = static.clone()
    _, offs = rt.grad_layout()
    for p, n, shp in zip(params, names, shapes):
        if p.grad is not None and lo <= p.grad.data_ptr() < hi:
            p.grad = keep[offs[n]:offs[n] + p.numel()].view(shp)


class _TrainStep(torch.autograd.Function):
    """forward = sedt_forward_train, backward = sedt_backward (hand-written kernels both ways); the trainable
    parameters are inputs so that autograd hands their gradients to the optimizer as usual (engine.py:70-80).
    Each forward takes its own runtime slot (tape + dropout stream), so several forwards may precede one backward
    (engine.py:134-170)."""

    @staticmethod
    def forward(ctx, model, x, mask, names, *params):
        rt = model.runtime(use_graph=model.use_cuda_graph)
        ctx.token = _Token()
        res, tctx = rt.forward_train(x, mask, use_graph=model.use_cuda_graph, dropout=float(model.transformer.dropout),
                                     token=ctx.token)
        ctx.model, ctx.tctx, ctx.names, ctx.shapes, ctx.params = model, tctx, names, [p.shape for p in params], params
        ctx.has_at = "at" in res
        outs = (res["logits"], res["boxes"]) + ((res["at"],) if ctx.has_at else ())
        if tctx.graph:
            # the slot's static output buffers are overwritten by its next replay; a second forward in flight uses another
            # slot, but outputs that outlive their backward (logging, EMA targets) must not alias them
            outs = tuple(o.clone() for o in outs)
        return outs

    @staticmethod
    def backward(ctx, d_logits, d_boxes, d_at=None):
        model = ctx.model
        rt = model._rt
        train_backbone = any(n.startswith("backbone.") for n in ctx.names)
        if ctx.tctx.graph:
            # graph mode: the gradients land in the slot's static buffer and views of it become p.grad.  If a p.grad still
            # aliases that buffer (gradient accumulation over micro-batches, engine.py:75-80, or zero_grad(set_to_none=False)),
            # the replay below would overwrite it and autograd would then add the buffer to itself: move those gradients to
            # their own storage first (one flat copy, only in that case).
            _detach_grads_from(ctx.tctx.slot.g.get("grads"), ctx.params, ctx.names, ctx.shapes, rt)
        post = _allreduce_buckets(rt, train_backbone, model.grad_allreduce_overlap) if model.grad_allreduce else None
        flat = rt.backward(ctx.tctx, d_logits, d_boxes, d_at, train_backbone, post=post)
        _, offs = rt.grad_layout()
        grads = []
        for n, shp in zip(ctx.names, ctx.shapes):
            k = 1
            for v in shp:
                k *= v
            grads.append(flat[offs[n]:offs[n] + k].view(shp))
        return (None, None, None, None, *grads)


class _TrainStepSP(torch.autograd.Function):
    """SP-SEDT pretraining step (sedt/spsedt.py:34-91 in train mode, frozen backbone): forward = sedt_forward_train_sp,
    backward = sedt_backward_sp.  Outputs: pred_logits, pred_boxes, pred_feature of all decoder layers (differentiable) and
    gt_feature (the frozen backbone's patch features: a constant)."""

    @staticmethod
    def forward(ctx, model, x, mask, patches, query_keep, names, *params):
        rt = model.runtime()
        ctx.token = _Token()
        res, tctx = rt.forward_train(x, mask, use_graph=False, dropout=float(model.transformer.dropout), token=ctx.token,
                                     patches=patches, query_keep=query_keep)
        ctx.model, ctx.tctx, ctx.names, ctx.shapes = model, tctx, names, [p.shape for p in params]
        ctx.has_feat = "pred_feature" in res
        outs = (res["logits"], res["boxes"], res["gt_feature"]) + ((res["pred_feature"],) if ctx.has_feat else ())
        ctx.mark_non_differentiable(res["gt_feature"])
        return outs

    @staticmethod
    def backward(ctx, d_logits, d_boxes, d_gt=None, d_feat=None):
        model = ctx.model
        rt = model._rt
        post = _allreduce_buckets(rt, False, False) if model.grad_allreduce else None
        flat = rt.backward(ctx.tctx, d_logits, d_boxes, None, False, d_pred_feature=d_feat, post=post)
        _, offs = rt.grad_layout()
        grads = []
        for n, shp in zip(ctx.names, ctx.shapes):
            k = 1
            for v in shp:
                k *= v
            grads.append(flat[offs[n]:offs[n] + k].view(shp))
        return (None, None, None, None, None, None, *grads)


class SEDT(nn.Module):
    """Drop-in for sedt.sedt.SEDT.  Extra keyword `precision`: "bf16" (default: bf16 operands,
    fp32 accumulation on tcgen05 tensor cores) or "fp32" (CUDA-core tier held to 1e-4 parity)."""

    def __init__(self, backbone, transformer, num_classes, num_queries, aux_loss=False, dec_at=False, pooling=None,
                 precision: str = "bf16", use_tensor_cores: bool = True):
        super().__init__()
        if pooling is not None:
            raise NotImplementedError("pooling/at_p heads (sedt/sedt.py:47-61,96-106) are unused by every documented "
                                      "recipe (default None) and are not part of the B200 hot path")
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self.num_queries = num_queries
        self.transformer = transformer
        hidden_dim = transformer.d_model
        self.class_embed = nn.Linear(hidden_dim, num_classes + 1)
        self.bbox_embed = MLP(hidden_dim, hidden_dim, 2, 3)
        self.input_proj = nn.Conv2d(backbone.num_channels, hidden_dim, kernel_size=1)
        self.backbone = backbone
        self.aux_loss = aux_loss
        self.dec_at = dec_at
        self.pooling = pooling
        self.num_classes = num_classes
        if self.dec_at:
            self.query_embed = nn.Embedding(num_queries + 1, hidden_dim)
            self.weak_class_embed = nn.Linear(hidden_dim, num_classes)
        else:
            self.query_embed = nn.Embedding(num_queries, hidden_dim)
        self.precision = precision
        self.use_tensor_cores = use_tensor_cores
        # replay the forward as one CUDA graph per input shape (outputs then live in runtime-owned buffers that
        # the next call with the same shape overwrites)
        self.use_cuda_graph = False
        # data-parallel training: average the flat gradient bucket over the ranks inside backward (the model is then
        # NOT wrapped in DistributedDataParallel; one all-reduce instead of DDP's per-bucket hooks)
        self.grad_allreduce = False
        # True: two buckets, the non-backbone one overlapped with the backbone backward (measured slower on B200, see
        # _allreduce_buckets); False: one all-reduce after the last backward kernel
        self.grad_allreduce_overlap = False
        self._rt: Optional[ForwardRuntime] = None
        self._self_sup = False
        self._feature_recon = False
        self._num_patches = 1

    # ---- native runtime plumbing ----------------------------------------------
    def _native_config(self) -> Dict[str, int]:
        tr = self.transformer
        lin1 = tr.encoder.layers[0].linear1 if len(tr.encoder.layers) else tr.decoder.layers[0].linear1
        return dict(enc_layers=len(tr.encoder.layers), dec_layers=len(tr.decoder.layers), num_queries=self.num_queries,
                    num_classes=self.num_classes, hidden_dim=tr.d_model, nheads=tr.nhead,
                    dim_feedforward=lin1.out_features, dec_at=int(self.dec_at), pre_norm=int(tr.normalize_before),
                    dilation=int(self._dilation),
                    self_sup=int(self._self_sup), feature_recon=int(self._feature_recon),
                    num_patches=int(self._num_patches), aux_loss=int(self.aux_loss),
                    precision=PRECISIONS[self.precision], use_tensor_cores=int(self.use_tensor_cores))

    @property
    def _dilation(self) -> bool:
        return bool(getattr(self.backbone[0], "dilation", True))

    def set_precision(self, precision: str, use_tensor_cores: bool = True) -> "SEDT":
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self.precision, self.use_tensor_cores = precision, use_tensor_cores
        self._rt = None
        return self

    def runtime(self, use_graph: bool = False) -> ForwardRuntime:
        if self._rt is None:
            self._rt = ForwardRuntime(self._native_config())
        tensors = dict(self.named_parameters())
        tensors.update(dict(self.named_buffers()))
        self._rt.ensure_packed(tensors, use_graph=use_graph)
        return self._rt

    def _wants_grad(self) -> bool:
        return torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters())

    def _train_unsupported(self) -> Optional[str]:
        """Why the native training kernels cannot run this module (None = they can)."""
        if self.precision != "bf16" or not self.use_tensor_cores:
            return "the training kernels exist for the bf16 tcgen05 tier only"
        if not self.transformer.normalize_before:
            return "the training kernels implement the pre-norm layers only"
        if not 0.0 <= float(self.transformer.dropout) < 1.0:
            return f"dropout must be in [0, 1), got {self.transformer.dropout}"
        return None

    def _check_mode(self):
        """Training with gradients runs through the native backward (bf16 tier, pre-norm); everything else that
        would need autograd raises instead of silently falling back."""
        if not self._wants_grad():
            return
        why = self._train_unsupported()
        if why is not None:
            raise NotImplementedError(why + ". Call model.eval() / torch.no_grad() for inference.")

    def _forward_train(self, x, mask):
        named = [(n, p) for n, p in self.named_parameters() if p.requires_grad]
        names = tuple(n for n, _ in named)
        outs = _TrainStep.apply(self, x, mask, names, *[p for _, p in named])
        res = {"logits": outs[0], "boxes": outs[1]}
        if self.dec_at:
            res["at"] = outs[2]
        return res

    def _device(self):
        return self.query_embed.weight.device

    def _prepare(self, samples):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("model parameters are on the CPU: call model.cuda() first (there is no CPU path)")
        if isinstance(samples, torch.Tensor) and samples.dim() == 4:
            # a dense [B,1,T,F] batch has no padding: skip building the all-False mask (utils.py:470-492)
            return samples.to(dev, torch.float32, non_blocking=True), None
        if isinstance(samples, (list, torch.Tensor)):
            samples = nested_tensor_from_tensor_list(samples)
        x, mask = samples.decompose()
        assert mask is not None
        x = x.to(dev, torch.float32, non_blocking=True)
        unpadded = getattr(samples, "unpadded", None)
        if unpadded is None:
            unpadded = not bool(mask.any().item())
        return x, (None if unpadded else mask.to(dev, non_blocking=True))

    # ---- forward ------------------------------------------------------------------
    def forward(self, samples: NestedTensor):
        """Same contract as sedt/sedt.py:64-123: returns pred_logits [B,Q,C+1], pred_boxes [B,Q,2]
        (center, width), `at` [B,C] under dec_at, and aux_outputs for the earlier decoder layers."""
        self._check_mode()
        x, mask = self._prepare(samples)
        if self._wants_grad():
            res = self._forward_train(x, mask)
        elif self.training and float(self.transformer.dropout) > 0.0 and self._train_unsupported() is None:
            # train() mode without gradients (the mean-teacher forward, engine.py:146-147): the reference keeps dropout
            # active there, so this runs the training forward (fresh Philox masks, nothing kept for a backward).  Where the
            # training kernels do not apply (fp32 tier, post-norm) the eval kernels run instead, i.e. dropout is off.
            res, _ = self.runtime().forward_train(x, mask, use_graph=False, dropout=float(self.transformer.dropout), token=None)
        else:
            res = self.runtime().forward(x, mask, use_graph=self.use_cuda_graph)
        out = {"pred_logits": res["logits"][-1], "pred_boxes": res["boxes"][-1]}
        if self.dec_at:
            out["at"] = res["at"].squeeze()              # sedt.py:92 squeezes: [C] when B == 1
        if self.aux_loss:
            out["aux_outputs"] = [{"pred_logits": a, "pred_boxes": b}
                                  for a, b in zip(res["logits"][:-1], res["boxes"][:-1])]
        return out


class SPSEDT(SEDT):
    """Drop-in for sedt.spsedt.SPSEDT: the test branch in eval(), the pretraining branch (random query drop, doubled query
    embedding, feature reconstruction head) with its backward in train()."""

    def __init__(self, backbone, transformer, num_classes, num_queries, aux_loss=False, dec_at=False, feature_recon=True,
                 query_shuffle=False, mask_ratio=0.1, num_patches=10, pooling=None, precision: str = "bf16",
                 use_tensor_cores: bool = True):
        super().__init__(backbone, transformer, num_classes, num_queries, aux_loss, dec_at, pooling, precision,
                         use_tensor_cores)
        if dec_at:
            raise NotImplementedError("SP-SEDT with an audio query is not a valid reference configuration "
                                      "(sedt/spsedt.py:59,72 shapes do not line up)")
        if query_shuffle:
            raise NotImplementedError("query_shuffle is off in every documented recipe (train_spsedt.py:42)")
        hidden_dim = transformer.d_model
        self.patch2query = nn.Linear(backbone.num_channels, hidden_dim)
        self.num_patches = num_patches
        self.mask_ratio = mask_ratio
        self.feature_recon = feature_recon
        if feature_recon:
            self.feature_align = MLP(hidden_dim, hidden_dim, backbone.num_channels, 2)
        self.query_shuffle = query_shuffle
        assert num_queries % num_patches == 0
        self._self_sup, self._feature_recon, self._num_patches = True, bool(feature_recon), num_patches

    def draw_query_keep(self, bs: int, device) -> torch.Tensor:
        """The reference's draw, verbatim (spsedt.py:65): torch.rand(num_queries, bs, 1) > mask_ratio, returned as [bs, Q] uint8.
        A run seeded like the reference draws the same mask."""
        m = (torch.rand(self.num_queries, bs, 1, device=device) > self.mask_ratio)
        return m[:, :, 0].t().contiguous().to(torch.uint8)

    def forward(self, samples, patches: torch.Tensor, query_keep: Optional[torch.Tensor] = None):
        """sedt/spsedt.py:34-91.  eval(): the test branch (any number of patches <= num_patches).  train(): the training branch
        (num_patches patches, random query drop, doubled query embedding, dropout) through the native training kernels --
        with gradients (loss.backward() fills .grad of every trainable parameter; the backbone is frozen, train_spsedt.py:50)
        or without (torch.no_grad()).  query_keep [B, Q] (1 = keep the patch feature) overrides the random draw (tests)."""
        if isinstance(samples, (list, tuple)) and len(samples) == 2 and torch.is_tensor(samples[0]) and samples[0].dim() == 4:
            samples = NestedTensor(samples[0], samples[1])          # engine.py:59 passes .decompose()
        x, mask = self._prepare(samples)
        if self.training:
            why = self._train_unsupported()
            if why is None and any(p.requires_grad for n, p in self.named_parameters() if n.startswith("backbone.")):
                why = "SP-SEDT pretraining keeps the backbone frozen (train_spsedt.py:50: lr_backbone = 0)"
            if why is None and int(patches.shape[1]) != self.num_patches:
                why = f"the training branch needs exactly num_patches = {self.num_patches} patches per clip (spsedt.py:63-69)"
            if why is not None:
                raise NotImplementedError(why + ". Call model.eval() for inference.")
            if query_keep is None:
                query_keep = self.draw_query_keep(x.shape[0], x.device)
            patches = patches.to(x.device, torch.float32)
            if torch.is_grad_enabled():
                named = [(n, p) for n, p in self.named_parameters() if p.requires_grad]
                outs = _TrainStepSP.apply(self, x, mask, patches, query_keep, tuple(n for n, _ in named), *[p for _, p in named])
                res = {"logits": outs[0], "boxes": outs[1], "gt_feature": outs[2]}
                if self.feature_recon:
                    res["pred_feature"] = outs[3]
            else:
                res, _ = self.runtime().forward_train(x, mask, dropout=float(self.transformer.dropout), token=None, patches=patches,
                                                      query_keep=query_keep)
        else:
            res = self.runtime().forward(x, mask, patches=patches, use_graph=self.use_cuda_graph)
        out = {"pred_logits": res["logits"][-1], "pred_boxes": res["boxes"][-1]}
        if self.feature_recon:
            out["pred_feature"] = res["pred_feature"][-1]
            out["gt_feature"] = res["gt_feature"]
            if self.aux_loss:
                out["aux_outputs"] = [{"pred_logits": a, "pred_boxes": b, "pred_feature": c, "gt_feature": res["gt_feature"]}
                                      for a, b, c in zip(res["logits"][:-1], res["boxes"][:-1], res["pred_feature"][:-1])]
        elif self.aux_loss:
            out["aux_outputs"] = [{"pred_logits": a, "pred_boxes": b}
                                  for a, b in zip(res["logits"][:-1], res["boxes"][:-1])]
        return out
