"""Parameter containers with the reference's module tree, so that
`state_dict()` names/shapes, the optimizer's "backbone in name" split
(train_sedt.py:234-240), the freeze policy (sedt/backbone.py:60-62), EMA over
named_parameters (utilities/utils.py:57-81) and checkpoint surgery
(train_sedt.py:243-266) keep working unchanged.

These modules only HOLD tensors.  They have no forward of their own: all
arithmetic of the hot path runs in the sm_100a kernels behind
libsedt_b200.so (see ../runtime.py), never in torch ops.
"""
from __future__ import annotations

import copy

import torch
from torch import nn

from ..spec import backbone_blocks


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard against accidental eager use
        raise RuntimeError(f"{type(self).__name__} only holds parameters; the forward runs in libsedt_b200.so")


class FrozenBatchNorm2d(_Holder):
    """Buffers of sedt/backbone.py:17-53 (statistics and affine terms are fixed);
    folded to scale/bias by sedt_model_pack."""

    def __init__(self, n: int):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        state_dict.pop(prefix + "num_batches_tracked", None)      # torchvision checkpoints carry it (backbone.py:33-41)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


def _conv(cin, cout, k, bias=False):
    m = nn.Conv2d(cin, cout, k, bias=bias)
    nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")    # torchvision resnet.py:208-210
    return m


class Bottleneck(_Holder):
    def __init__(self, blk):
        super().__init__()
        self.conv1 = _conv(blk.conv1.cin, blk.conv1.cout, 1)
        self.bn1 = FrozenBatchNorm2d(blk.conv1.cout)
        self.conv2 = _conv(blk.conv2.cin, blk.conv2.cout, 3)
        self.bn2 = FrozenBatchNorm2d(blk.conv2.cout)
        self.conv3 = _conv(blk.conv3.cin, blk.conv3.cout, 1)
        self.bn3 = FrozenBatchNorm2d(blk.conv3.cout)
        if blk.downsample is not None:
            self.downsample = nn.Sequential(_conv(blk.downsample.cin, blk.downsample.cout, 1),
                                            FrozenBatchNorm2d(blk.downsample.cout))


class Body(_Holder):
    """conv0 + ResNet-50 trunk (sedt/backbone.py:97-111; IntermediateLayerGetter drops maxpool_/fc)."""

    def __init__(self, dilation: bool):
        super().__init__()
        self.conv0 = nn.Conv2d(1, 3, 1)
        self.conv1 = _conv(3, 64, 7)
        self.bn1 = FrozenBatchNorm2d(64)
        layers = {}
        for blk in backbone_blocks(dilation):
            li = blk.prefix.split("layer")[1].split(".")[0]
            layers.setdefault(li, []).append(Bottleneck(blk))
        for li, blocks in layers.items():
            setattr(self, f"layer{li}", nn.Sequential(*blocks))


class Backbone(_Holder):
    def __init__(self, name: str, train_backbone: bool, dilation: bool):
        super().__init__()
        if name != "resnet50":
            raise NotImplementedError(f"backbone '{name}': the B200 path implements resnet50, the only backbone of the "
                                      "documented recipes (train_sedt.py:74)")
        self.body = Body(dilation)
        self.dilation = bool(dilation)
        self.num_channels = 2048
        for pname, p in self.body.named_parameters():          # sedt/backbone.py:60-62
            if not train_backbone or ("conv0" not in pname and "layer2" not in pname and "layer3" not in pname
                                      and "layer4" not in pname):
                p.requires_grad_(False)


class PositionEmbeddingSine(_Holder):
    """No parameters; the table is built on the device (sedt/position_encoding.py:11-47)."""

    def __init__(self, num_pos_feats=256, temperature=10000, normalize=True):
        super().__init__()
        self.num_pos_feats, self.temperature, self.normalize = num_pos_feats, temperature, normalize


class Joiner(nn.Sequential):
    def __init__(self, backbone, position_embedding):
        super().__init__(backbone, position_embedding)
        self.num_channels = backbone.num_channels

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("Joiner only holds parameters; the forward runs in libsedt_b200.so")


class EncoderLayer(_Holder):
    def __init__(self, d, nhead, ff, dropout):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d, ff)
        self.linear2 = nn.Linear(ff, d)
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = nn.LayerNorm(d)


class DecoderLayer(_Holder):
    def __init__(self, d, nhead, ff, dropout):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, nhead, dropout=dropout)
        self.multihead_attn = nn.MultiheadAttention(d, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d, ff)
        self.linear2 = nn.Linear(ff, d)
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = nn.LayerNorm(d)
        self.norm3 = nn.LayerNorm(d)


class _Stack(_Holder):
    def __init__(self, layer, n, norm):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(layer) for _ in range(n)])    # transformer.py:405-406
        self.num_layers = n
        self.norm = norm


class Transformer(_Holder):
    """Parameter tree of sedt/transformer.py:17-46."""

    def __init__(self, d_model=256, nhead=8, num_encoder_layers=3, num_decoder_layers=3, dim_feedforward=2048,
                 dropout=0.1, normalize_before=True, self_sup=False):
        super().__init__()
        self.encoder = _Stack(EncoderLayer(d_model, nhead, dim_feedforward, dropout), num_encoder_layers,
                              nn.LayerNorm(d_model) if normalize_before else None)
        self.decoder = _Stack(DecoderLayer(d_model, nhead, dim_feedforward, dropout), num_decoder_layers,
                              nn.LayerNorm(d_model))
        for p in self.parameters():                              # transformer.py:42-45
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self.d_model, self.nhead, self.self_sup = d_model, nhead, self_sup
        self.normalize_before, self.dropout = normalize_before, dropout


class MLP(_Holder):
    """Parameter tree of sedt/sedt.py:398-409."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))
