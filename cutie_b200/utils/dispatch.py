"""Deterministic dispatch of the pixel-side glue ops inside the PyTorch stages around the hot path.

The glue kernels in csrc/pixel.cu replace short chains of ATen launches (area down-sampling, the CAResBlock tail, the
sensory GRU gates, the end of CUTIE.segment, the decoder's one-channel prediction head).  Which form runs is read from
a COMMITTED table keyed by op name -- the same video executes the same arithmetic on every run, rank and GPU.  (Round 1
chose the form by an on-device A/B at first use; two runs could then execute different arithmetic.  The table below is
what those A/Bs measured on B200 with fp32 convolutions: profiles/r02_*.)  CPU tensors (the oracle harness borrowing the
modules) always take the ATen form; a `KernelError` from our own library propagates like everywhere else in cutie_b200.

Attached per model by `CUTIE.optimize_for_inference()` (attribute `glue_dispatch` on every sub-module); modules without
it run PyTorch's launches unchanged.
"""
from typing import Callable, Dict

import torch
import torch.nn as nn

# op -> True: the cutie_b200 kernel, False: the ATen composition
GLUE_TABLE: Dict[str, bool] = {
    'area_pool': True,                    # cutie_area_pool vs adaptive_avg_pool (15 CTAs at 480p)
    'eca_scale_add': True,                # conv1d + sigmoid + mul + add -> one stream
    'gated_update': True,                 # 8 launches -> 1
    'segment_tail': True,                 # sigmoid, aggregate, bilinear x4, softmax: 11 launches -> 2
    'pred_conv3x3': True,                 # one-output-channel 3x3 head: one pass over the 40 MB input
    'caresblock_channels_last': False,    # channels-last PixelFFN twin: slower than NCHW with fp32 (non-TF32) engines
    'objresblock_channels_last': False,   # same for the decoder's residual blocks
}


class GlueDispatch:
    def __init__(self, enabled: bool = True, table: Dict[str, bool] = None):
        self.enabled = enabled
        self.table = dict(GLUE_TABLE if table is None else table)
        self.calls: Dict[str, int] = {}

    def _eligible(self, probe: torch.Tensor) -> bool:
        return self.enabled and probe.is_cuda and probe.dtype == torch.float32 and not torch.is_grad_enabled()

    def __call__(self, op: str, key: tuple, aten: Callable, kernel: Callable, probe: torch.Tensor, rtol: float = 0.0):
        """aten(): the PyTorch composition; kernel(trial: bool): our kernel (always called with trial=False here)."""
        if not self._eligible(probe) or not self.table.get(op, False):
            return aten()
        self.calls[op] = self.calls.get(op, 0) + 1
        return kernel(False)

    def __deepcopy__(self, memo):
        new = GlueDispatch(self.enabled, self.table)
        memo[id(self)] = new
        return new

    def report(self) -> dict:
        return {'enabled': self.enabled, 'table': dict(self.table)}


def attach_glue_dispatch(module: nn.Module, dispatch: GlueDispatch) -> int:
    n = 0
    for m in module.modules():
        object.__setattr__(m, 'glue_dispatch', dispatch)
        n += 1
    return n
