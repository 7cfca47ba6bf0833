"""Padding / soft-aggregation helpers used on the path (cutie/utils/tensor_utils.py:7-54)."""
from typing import Iterable, Tuple

import torch
import torch.nn.functional as F


def pad_divide_by(in_img: torch.Tensor, d: int) -> Tuple[torch.Tensor, Iterable[int]]:
    """Zero-pad the last two dims up to multiples of d, split as evenly as possible (extra on the far side)."""
    h, w = in_img.shape[-2:]
    new_h, new_w = -(-h // d) * d, -(-w // d) * d
    lh, lw = (new_h - h) // 2, (new_w - w) // 2
    pad = (lw, new_w - w - lw, lh, new_h - h - lh)
    return F.pad(in_img, pad), pad


def unpad(img: torch.Tensor, pad: Iterable[int]) -> torch.Tensor:
    lw, uw, lh, uh = pad
    if img.dim() not in (3, 4, 5):
        raise NotImplementedError
    H, W = img.shape[-2:]
    return img[..., lh:H - uh, lw:W - uw]


def aggregate(prob: torch.Tensor, dim: int) -> torch.Tensor:
    """Soft aggregation: prepend background = prod(1-p), clamp, return log-odds (tensor_utils.py:47-54)."""
    prob = prob.float()
    bg = torch.prod(1 - prob, dim=dim, keepdim=True)
    p = torch.cat([bg, prob], dim).clamp(1e-7, 1 - 1e-7)
    return torch.log(p / (1 - p))
