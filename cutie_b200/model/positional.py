"""2-D sinusoidal positional encoding (cutie/model/transformer/positional_encoding.py:20-97).

One [h, w, E] map per spatial size, identical for every object; computed once per shape (the reference
also caches, :64-68) so it never appears on the per-frame path.
"""
import math

import torch
import torch.nn as nn


class SinusoidPE(nn.Module):
    def __init__(self, embed_dim: int, scale: float, temperature: float):
        super().__init__()
        self.half = int(math.ceil(embed_dim / 4) * 2)
        self.scale = scale
        self.register_buffer('inv_freq',
                             1.0 / (temperature ** (torch.arange(0, self.half, 2).float() / self.half)))
        self._cache = {}

    def grid(self, h: int, w: int) -> torch.Tensor:
        """[h, w, 2*half]: channels [0,half) encode x, [half,2*half) encode y; sin/cos interleaved."""
        key = (h, w, self.inv_freq.device)
        if key not in self._cache:
            f = self.inv_freq

            def axis(n):
                p = torch.arange(n, device=f.device, dtype=f.dtype)
                p = p / (p[-1] + 1e-6) * self.scale
                a = p[:, None] * f[None]
                return torch.stack((a.sin(), a.cos()), -1).flatten(-2)
            ex, ey = axis(w), axis(h)
            self._cache = {key: torch.cat([ex[None].expand(h, -1, -1), ey[:, None].expand(-1, w, -1)], -1)
                           .contiguous()}
        return self._cache[key]
