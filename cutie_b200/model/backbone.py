"""ResNet-18/50 trunks truncated after stage 3 (stride 16) -- cuDNN convolutions, kept as PyTorch calls
(BASELINE.json north_star: "the encoder/decoder convolutions stay as PyTorch/cuDNN calls").

Parameter names follow torchvision-style ResNets as the reference stores them
(cutie/model/utils/resnet.py:54-164) so `cutie-base-mega.pth` loads: conv1/bn1, layerN.i.{conv,bn}{1,2,3},
layerN.0.downsample.{0,1}.  BatchNorm always runs on its running statistics (inference only).
"""
from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from cutie_b200.model.fuse import conv_add_relu, conv_relu


class _Residual(nn.Module):
    """One residual unit; `widths` lists the conv output widths, `ksizes` their kernel sizes."""

    def __init__(self, c_in: int, widths: Sequence[int], ksizes: Sequence[int], stride: int):
        super().__init__()
        c = c_in
        stride_at = 0 if len(widths) == 2 else 1      # basic: first conv strides; bottleneck: the 3x3
        for i, (wd, ks) in enumerate(zip(widths, ksizes), start=1):
            s = stride if (i - 1) == stride_at else 1
            setattr(self, f'conv{i}', nn.Conv2d(c, wd, ks, stride=s, padding=ks // 2, bias=False))
            setattr(self, f'bn{i}', nn.BatchNorm2d(wd))
            c = wd
        self.n = len(widths)
        if stride != 1 or c_in != c:
            self.downsample = nn.Sequential(nn.Conv2d(c_in, c, 1, stride=stride, bias=False), nn.BatchNorm2d(c))
        else:
            self.downsample = None

    def forward(self, x):
        if getattr(self, 'bn_folded', False):
            # BatchNorms folded into the convolutions (fuse.fold_trunk_): every ReLU and the residual add ride in
            # the convolution's epilogue where cuDNN's fused graph wins its on-device trial (fuse.ConvEpilogueFuser)
            y = x
            for i in range(1, self.n):
                y = conv_relu(getattr(self, f'conv{i}'), y)
            skip = x if self.downsample is None else self.downsample(x)
            return conv_add_relu(getattr(self, f'conv{self.n}'), y, skip)
        y = x
        for i in range(1, self.n + 1):
            y = getattr(self, f'bn{i}')(getattr(self, f'conv{i}')(y))
            if i < self.n:
                y = F.relu(y, inplace=True)
        skip = x if self.downsample is None else self.downsample(x)
        return F.relu(y + skip, inplace=True)


def _stage(kind: str, c_in: int, planes: int, count: int, stride: int) -> nn.Sequential:
    if kind == 'basic':
        widths, ks, c_out = (planes, planes), (3, 3), planes
    else:
        widths, ks, c_out = (planes, planes, planes * 4), (1, 3, 1), planes * 4
    units = [_Residual(c_in, widths, ks, stride)]
    units += [_Residual(c_out, widths, ks, 1) for _ in range(count - 1)]
    return nn.Sequential(*units)


class ResNetTrunk(nn.Module):
    """Stem + stages 1..3.  `extra_in` adds input channels to the stem (mask + others for the mask encoder)."""
    SPECS = {'resnet18': ('basic', (2, 2, 2), 1), 'resnet50': ('bottleneck', (3, 4, 6), 4)}

    def __init__(self, arch: str, extra_in: int = 0):
        super().__init__()
        if arch not in self.SPECS:
            raise NotImplementedError(arch)
        kind, counts, exp = self.SPECS[arch]
        self.conv1 = nn.Conv2d(3 + extra_in, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.layer1 = _stage(kind, 64, 64, counts[0], 1)
        self.layer2 = _stage(kind, 64 * exp, 128, counts[1], 2)
        self.layer3 = _stage(kind, 128 * exp, 256, counts[2], 2)
