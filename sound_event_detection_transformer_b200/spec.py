"""Static description of the SEDT hot path: which convolutions, in which order,
with which geometry.  One table, shared by the weight packer, the synthetic
weight generator, the oracle and the tests, so they cannot drift apart.

Reference for the geometry: sedt/backbone.py:89-113 (conv0 + torchvision
resnet50, replace_stride_with_dilation=[False, False, dilation]) and
torchvision/models/resnet.py:108-163,225-262 (Bottleneck v1.5: stride on the
3x3; layer4 with dilation keeps stride 1, block 0 uses the previous dilation
1, blocks 1-2 use dilation 2).
"""
from __future__ import annotations

import argparse
from dataclasses import dataclass
from typing import List, Optional

BODY = "backbone.0.body."

# (planes, blocks, stride) per ResNet-50 stage
_STAGES = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))


@dataclass(frozen=True)
class ConvSpec:
    """One convolution + FrozenBatchNorm pair of the backbone."""
    name: str          # state_dict prefix of the conv weight (without ".weight")
    bn: str            # state_dict prefix of the FrozenBatchNorm2d buffers
    cin: int
    cout: int
    k: int             # square kernel size (1, 3 or 7)
    stride: int
    dilation: int
    pad: int
    relu: bool         # ReLU directly after the BN (before any residual add)


@dataclass(frozen=True)
class BlockSpec:
    """One Bottleneck: conv1(1x1) -> conv2(3x3) -> conv3(1x1) (+ downsample)."""
    prefix: str
    conv1: ConvSpec
    conv2: ConvSpec
    conv3: ConvSpec
    downsample: Optional[ConvSpec]


def backbone_blocks(dilation: bool = True) -> List[BlockSpec]:
    blocks: List[BlockSpec] = []
    inplanes = 64
    cur_dil = 1
    for li, (planes, nblocks, stride) in enumerate(_STAGES, start=1):
        prev_dil = cur_dil
        if li == 4 and dilation:
            cur_dil *= stride
            stride = 1
        for bi in range(nblocks):
            p = f"{BODY}layer{li}.{bi}."
            s = stride if bi == 0 else 1
            d = prev_dil if bi == 0 else cur_dil
            cin = inplanes
            ds = None
            if bi == 0 and (s != 1 or cin != planes * 4):
                ds = ConvSpec(p + "downsample.0", p + "downsample.1", cin, planes * 4, 1, s, 1, 0, False)
            blocks.append(BlockSpec(
                prefix=p,
                conv1=ConvSpec(p + "conv1", p + "bn1", cin, planes, 1, 1, 1, 0, True),
                conv2=ConvSpec(p + "conv2", p + "bn2", planes, planes, 3, s, d, d, True),
                conv3=ConvSpec(p + "conv3", p + "bn3", planes, planes * 4, 1, 1, 1, 0, False),
                downsample=ds))
            inplanes = planes * 4
    return blocks


def conv_out(n: int, k: int, stride: int, pad: int, dil: int = 1) -> int:
    return (n + 2 * pad - dil * (k - 1) - 1) // stride + 1


def feature_hw(T: int, F: int, dilation: bool = True):
    """Spatial size after each stage: returns [(H,W) for stem, layer1..layer4]."""
    h, w = conv_out(T, 7, 2, 3), conv_out(F, 7, 2, 3)      # conv1
    h, w = conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)      # maxpool
    out = [(h, w)]
    for li, (_, _, stride) in enumerate(_STAGES, start=1):
        if li == 4 and dilation:
            stride = 1
        if stride == 2:
            h, w = conv_out(h, 3, 2, 1), conv_out(w, 3, 2, 1)
        out.append((h, w))
    return out


def default_args(**overrides) -> argparse.Namespace:
    """The argparse namespace `build_model(args)` reads (sedt/__init__.py:8-63,
    train_sedt.py:28-129 defaults).  Config-1 defaults; override for others."""
    d = dict(
        self_sup=False, num_classes=10, num_queries=10, aux_loss=True, dec_at=True, pooling=None,
        feature_recon=False, query_shuffle=False, num_patches=10,
        ce_loss_coef=1.0, bbox_loss_coef=5.0, giou_loss_coef=2.0, weak_loss_coef=1.0, weak_loss_p_coef=1.0,
        dec_layers=3, enc_layers=3, eos_coef=0.1, hidden_dim=256, position_embedding="sine",
        lr_backbone=1e-4, backbone="resnet50", dilation=True, dropout=0.1, nheads=8,
        dim_feedforward=2048, pre_norm=True,
        set_cost_class=1.0, set_cost_bbox=5.0, set_cost_giou=2.0, epsilon=1.0, alpha=1.0,
    )
    d.update(overrides)
    return argparse.Namespace(**d)


# The named workloads of BASELINE.json `configs`.
def config_args(name: str) -> argparse.Namespace:
    if name == "c1":    # SEDT E=3, URBAN-SED shape
        return default_args(enc_layers=3, num_queries=10, num_classes=10)
    if name == "c2":    # SEDT E=6, DCASE shape (train_sedt.py:151-152 forces 20 queries)
        return default_args(enc_layers=6, num_queries=20, num_classes=10)
    if name == "c5":    # SP-SEDT pretraining forward
        return default_args(enc_layers=6, num_queries=20, num_classes=1, self_sup=True, dec_at=False,
                            feature_recon=True, num_patches=10, lr_backbone=0.0)
    raise KeyError(name)


CONFIG_INPUT = {"c1": dict(B=64, T=500, F=64), "c2": dict(B=256, T=496, F=64),
                "c5": dict(B=200, T=496, F=64, P=10, PT=128)}
