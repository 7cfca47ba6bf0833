"""Inference-time graph surgery for the PyTorch/cuDNN stages around the hot path (they stay PyTorch calls):
eval-mode BatchNorm folded into the preceding convolution, ResNet trunks run channels-last so cuDNN's NHWC
tensor-core kernels need no per-layer NCHW<->NHWC conversion.  Applied AFTER weights are loaded
(`CUTIE.optimize_for_inference()`); the state_dict layout of an optimised model is no longer the checkpoint's."""
import torch
import torch.nn as nn


def fold_conv_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d) -> nn.Conv2d:
    scale = bn.weight.detach() / torch.sqrt(bn.running_var.detach() + bn.eps)
    fused = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding,
                      conv.dilation, conv.groups, bias=True).to(conv.weight.device, conv.weight.dtype)
    fused.weight.data = conv.weight.detach() * scale.view(-1, 1, 1, 1)
    bias = bn.bias.detach() - bn.running_mean.detach() * scale
    if conv.bias is not None:
        bias = bias + conv.bias.detach() * scale
    fused.bias.data = bias
    return fused


def fold_trunk_(module: nn.Module) -> int:
    """Folds every (convN, bnN) pair and (downsample.0, downsample.1) pair found under `module`, in place."""
    n = 0
    for m in module.modules():
        for i in (1, 2, 3):
            conv, bn = getattr(m, f'conv{i}', None), getattr(m, f'bn{i}', None)
            if isinstance(conv, nn.Conv2d) and isinstance(bn, nn.BatchNorm2d):
                setattr(m, f'conv{i}', fold_conv_bn(conv, bn))
                setattr(m, f'bn{i}', nn.Identity())
                n += 1
        ds = getattr(m, 'downsample', None)
        if isinstance(ds, nn.Sequential) and len(ds) == 2 and isinstance(ds[1], nn.BatchNorm2d):
            m.downsample = nn.Sequential(fold_conv_bn(ds[0], ds[1]), nn.Identity())
            n += 1
    return n
