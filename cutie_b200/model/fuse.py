"""Inference-time graph surgery for the PyTorch/cuDNN stages around the hot path (they stay PyTorch calls):
eval-mode BatchNorm folded into the preceding convolution, ResNet trunks run channels-last so cuDNN's NHWC
tensor-core kernels need no per-layer NCHW<->NHWC conversion.  Applied AFTER weights are loaded
(`CUTIE.optimize_for_inference()`); the state_dict layout of an optimised model is no longer the checkpoint's."""
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401


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
        if all(isinstance(getattr(m, f'bn{i}', nn.Identity()), nn.Identity) for i in (1, 2, 3)) and \
                isinstance(getattr(m, 'conv1', None), nn.Conv2d):
            m.bn_folded = True          # _Residual / encoder stems switch to the fused-epilogue forward
    return n


# ---------------------------------------------------------------------------------------------------
# conv + bias (+ residual) + ReLU as ONE cuDNN call
# ---------------------------------------------------------------------------------------------------
class ConvEpilogueFuser:
    """`relu(conv(x) + bias [+ z])` through cuDNN's fused conv-bias-add-activation graph
    (`torch.cudnn_convolution_relu` / `torch.cudnn_convolution_add_relu`, the ops PyTorch's own frozen-graph
    pass emits) instead of three launches (convolution, broadcast bias add, clamp).

    On the round-1 launch list the trunks' bias adds and ReLUs were ~150 launches and ~1 ms of a 4.35 ms frame
    (`profiles/r01_ncu_summary.md`: `elementwise_kernel<add>` 85/step at 8 us, `clamp_scalar` 62/step at 4.8 us,
    residual adds 26/step at 6.7 us).  These stay PyTorch/cuDNN calls -- only the call changes.

    Nothing is assumed about how the fused engines behave on a given GPU / cuDNN build: the first time a
    (layer, input geometry) pair is seen OUTSIDE a stream capture, both forms run on the live tensors, the fused
    result must match the three-launch result, both are timed with CUDA events, and the faster one is kept
    for that pair (`decisions`).  Any exception from the fused op keeps the three-launch form and is recorded in
    `errors`.  CPU tensors (the oracle harness borrowing these modules) always take the three-launch form.

    `cudnn_convolution_relu` hands the *uninitialised* output to cuDNN as the residual operand with alpha = 0;
    0 x (stale NaN bits) is NaN, so the no-residual case passes a persistent zero tensor of the output shape
    instead (read once per call, ~100 MB per 480p frame over all layers: 15 us of HBM time).
    """

    def __init__(self, enabled: bool = True, trial_iters: int = 6):
        self.enabled = enabled
        self.trial_iters = trial_iters
        self.decisions = {}          # key -> True (fused) / False (three launches)
        self.timings = {}            # key -> (fused_ms, unfused_ms)
        self.errors = []
        self._zeros = {}

    # -- the two forms --------------------------------------------------------------------------------
    @staticmethod
    def unfused(conv: nn.Conv2d, x: torch.Tensor, z=None) -> torch.Tensor:
        # nn.Conv2d's own convolution (not conv(x): ObjConv2d overrides forward for 5-D object tensors)
        y = conv._conv_forward(x, conv.weight, conv.bias)
        if z is not None:
            y = y.add_(z) if not y.requires_grad else y + z
        return torch.relu_(y) if not y.requires_grad else torch.relu(y)

    def _zero_like_output(self, conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
        n, _, h, w = x.shape
        ho = (h + 2 * conv.padding[0] - conv.dilation[0] * (conv.kernel_size[0] - 1) - 1) // conv.stride[0] + 1
        wo = (w + 2 * conv.padding[1] - conv.dilation[1] * (conv.kernel_size[1] - 1) - 1) // conv.stride[1] + 1
        cl = x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
        cl = cl or conv.weight.is_contiguous(memory_format=torch.channels_last) and not conv.weight.is_contiguous()
        key = (n, conv.out_channels, ho, wo, x.dtype, x.device, bool(cl))
        buf = self._zeros.get(key)
        if buf is None:
            buf = torch.empty(n, conv.out_channels, ho, wo, dtype=x.dtype, device=x.device,
                              memory_format=torch.channels_last if cl else torch.contiguous_format).zero_()
            self._zeros[key] = buf
        return buf

    def fused(self, conv: nn.Conv2d, x: torch.Tensor, z=None) -> torch.Tensor:
        if z is None:
            return torch.cudnn_convolution_add_relu(x, conv.weight, self._zero_like_output(conv, x), 0.0, conv.bias,
                                                    conv.stride, conv.padding, conv.dilation, conv.groups)
        return torch.cudnn_convolution_add_relu(x, conv.weight, z, 1.0, conv.bias,
                                                conv.stride, conv.padding, conv.dilation, conv.groups)

    # -- one-off trial per (layer, geometry) -------------------------------------------------------------
    @staticmethod
    def _key(conv, x, z):
        return (id(conv), tuple(x.shape), tuple(x.stride()), x.dtype, z is not None,
                torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)

    def _time(self, fn) -> float:
        for _ in range(2):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(self.trial_iters):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / self.trial_iters

    def _trial(self, key, conv, x, z) -> bool:
        try:
            # the three-launch form adds z in place into a fresh tensor, never into z itself
            ref = self.unfused(conv, x, z)
            out = self.fused(conv, x, z)
            scale = float(ref.abs().max()) + 1e-6
            err = float((out - ref).abs().max())
            tol = (2e-2 if torch.backends.cudnn.allow_tf32 else 2e-4) * scale
            if not (err <= tol):                   # also catches NaN
                self.errors.append(f'conv {tuple(conv.weight.shape)} on {tuple(x.shape)}: fused differs by {err:.3e} '
                                   f'(scale {scale:.3e})')
                return False
            t_f = self._time(lambda: self.fused(conv, x, z))
            t_u = self._time(lambda: self.unfused(conv, x, z))
            self.timings[key] = (t_f, t_u)
            return t_f <= t_u
        except Exception as e:                     # noqa: BLE001 -- any cuDNN / dispatcher failure: keep three launches
            self.errors.append(f'conv {tuple(conv.weight.shape)} on {tuple(x.shape)}: {type(e).__name__}: {e}')
            return False

    def _eligible(self, conv: nn.Conv2d, x: torch.Tensor) -> bool:
        return (self.enabled and x.is_cuda and conv.bias is not None and conv.padding_mode == 'zeros'
                and x.dim() == 4 and not torch.is_grad_enabled())

    @staticmethod
    def _capturing() -> bool:
        return torch.cuda.is_current_stream_capturing()

    def __call__(self, conv: nn.Conv2d, x: torch.Tensor, z=None) -> torch.Tensor:
        if not self._eligible(conv, x):
            return self.unfused(conv, x, z)
        key = self._key(conv, x, z)
        use = self.decisions.get(key)
        if use is None:
            if self._capturing():
                return self.unfused(conv, x, z)    # no timing inside a capture; _Captured warms up outside one first
            use = self.decisions[key] = self._trial(key, conv, x, z)
        return self.fused(conv, x, z) if use else self.unfused(conv, x, z)

    def report(self) -> dict:
        n_f = sum(1 for v in self.decisions.values() if v)
        saved = sum(u - f for k, (f, u) in self.timings.items() if self.decisions.get(k))
        return {'enabled': self.enabled, 'fused': n_f, 'three_launch': len(self.decisions) - n_f,
                'errors': len(self.errors), 'first_error': self.errors[0] if self.errors else None,
                'trial_ms_saved_per_pass': saved}


def attach_epilogue_fuser(module: nn.Module, fuser: 'ConvEpilogueFuser') -> int:
    """Hands `fuser` to every nn.Conv2d under `module` (plain attribute, not a parameter / buffer / sub-module, so
    state_dict and .to() are unaffected).  Per model, not process-wide: an un-optimised model keeps three launches."""
    n = 0
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            object.__setattr__(m, 'epilogue_fuser', fuser)
            n += 1
    return n


def conv_relu(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """relu(conv(x)) -- one cuDNN call where the model's fuser says so, else convolution + bias add + clamp."""
    f = getattr(conv, 'epilogue_fuser', None)
    return ConvEpilogueFuser.unfused(conv, x) if f is None else f(conv, x)


def conv_add_relu(conv: nn.Conv2d, x: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """relu(conv(x) + z)."""
    f = getattr(conv, 'epilogue_fuser', None)
    return ConvEpilogueFuser.unfused(conv, x, z) if f is None else f(conv, x, z)
