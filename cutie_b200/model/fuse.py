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
# convolution epilogues: bias (+ residual) (+ ReLU) without ATen's extra launches
# ---------------------------------------------------------------------------------------------------
class ConvEpilogueFuser:
    """`act(conv(x) + bias [+ z])` in fewer launches than PyTorch's convolution + broadcast bias add [+ add] [+ clamp].

    On the round-1 launch list those epilogue launches were ~180 per 4.35 ms frame and ~1.3 ms of it
    (`profiles/r01_ncu_summary.md`: `elementwise_kernel<add>` 85/step at 8 us -- the bias adds, which ATen runs through
    a broadcast TensorIterator with no vector accesses --, `clamp_scalar` 62/step at 4.8 us, residual adds 26/step at
    6.7 us).  The convolutions themselves stay cuDNN calls (BASELINE.json north_star); three forms of the epilogue:

      'aten'    F.conv2d(x, w, b) [+ y.add_(z)] [+ relu_]                -- what PyTorch does; the reference form
      'cudnn'   torch.cudnn_convolution_add_relu(x, w, z, alpha, b, ..)  -- cuDNN's fused conv-bias-add-ReLU graph (the
                op PyTorch's own frozen-graph pass emits); ReLU epilogues only
      'kernel'  F.conv2d(x, w, None) + cutie_bias_act(y, b, z, relu)     -- the bias-less convolution followed by ONE
                float4 stream of ours (csrc/pixel.cu), same association as 'aten' => bit-identical results

    Convolutions with a handful of input channels (the two ResNet stems: 3 and 5) get every form a second time as
    `<form>+pad`: the input is zero-padded to a multiple of 4 channels and the weight likewise (cached twin), which is
    what cuDNN's NHWC tensor-core kernels need for 16-byte channel vectors -- with C = 3 the round-1 profile shows the
    generic `convolve_common_engine_float_NHWC` at 141 us per stem call.  The zero channel contributes exact zeros.

    Nothing is assumed about how any form behaves on a given GPU / cuDNN build: the first time a (layer, input geometry,
    epilogue) triple is seen OUTSIDE a stream capture, every applicable form runs on the live tensors, a candidate must
    match 'aten', all are timed with CUDA events, and the fastest is kept for that triple (`decisions`, `timings`).
    A form that raises is dropped for that triple and recorded in `errors`.  CPU tensors (the oracle harness borrowing
    these modules) always take 'aten'.

    `cudnn_convolution_relu` hands the *uninitialised* output to cuDNN as the residual operand with alpha = 0;
    0 x (stale NaN bits) is NaN, so the no-residual case passes a persistent zero tensor of the output shape
    instead (read once per call, ~100 MB per 480p frame over all layers: 15 us of HBM time).
    """
    FORMS = ('aten', 'cudnn', 'kernel')

    def __init__(self, enabled: bool = True, trial_iters: int = 6, forms=FORMS):
        self.enabled = enabled
        self.trial_iters = trial_iters
        self.forms = tuple(forms)
        self.decisions = {}          # key -> form name
        self.timings = {}            # key -> {form: ms}
        self.errors = []
        self._zeros = {}
        self._twins = {}             # id(conv) -> (conv, twin with zero-padded input channels)

    # -- the forms ------------------------------------------------------------------------------------
    @staticmethod
    def _conv(conv: nn.Conv2d, x: torch.Tensor, with_bias: bool) -> torch.Tensor:
        # nn.Conv2d's own convolution (class method: instances may carry a patched _conv_forward, and ObjConv2d
        # overrides forward for 5-D object tensors)
        return nn.Conv2d._conv_forward(conv, x, conv.weight, conv.bias if with_bias else None)

    @classmethod
    def unfused(cls, conv: nn.Conv2d, x: torch.Tensor, z=None, relu: bool = True) -> torch.Tensor:
        """The 'aten' form."""
        y = cls._conv(conv, x, True)
        if z is not None:
            y = y.add_(z) if not y.requires_grad else y + z
        if relu:
            y = torch.relu_(y) if not y.requires_grad else torch.relu(y)
        return y

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
        """The 'cudnn' form (always with ReLU)."""
        if z is None:
            return torch.cudnn_convolution_add_relu(x, conv.weight, self._zero_like_output(conv, x), 0.0, conv.bias,
                                                    conv.stride, conv.padding, conv.dilation, conv.groups)
        return torch.cudnn_convolution_add_relu(x, conv.weight, z, 1.0, conv.bias,
                                                conv.stride, conv.padding, conv.dilation, conv.groups)

    def kernel(self, conv: nn.Conv2d, x: torch.Tensor, z=None, relu: bool = True) -> torch.Tensor:
        """The 'kernel' form: bias-less convolution + cutie_bias_act."""
        from cutie_b200 import kernels as K_
        return K_.bias_act_(self._conv(conv, x, False), conv.bias, z, relu)

    @staticmethod
    def _padded_channels(conv: nn.Conv2d) -> int:
        """Input channels after zero padding, or 0 if this convolution does not get '+pad' forms."""
        c = conv.in_channels
        if conv.groups != 1 or c % 4 == 0 or c > 12:
            return 0
        return (c + 3) // 4 * 4

    def _twin(self, conv: nn.Conv2d) -> nn.Conv2d:
        got = self._twins.get(id(conv))
        if got is not None and got[0] is conv and got[1].weight.device == conv.weight.device:
            return got[1]
        cp = self._padded_channels(conv)
        tw = nn.Conv2d(cp, conv.out_channels, conv.kernel_size, conv.stride, conv.padding, conv.dilation, conv.groups,
                       bias=conv.bias is not None)
        w = torch.zeros(conv.out_channels, cp, *conv.kernel_size, dtype=conv.weight.dtype, device=conv.weight.device)
        w[:, :conv.in_channels] = conv.weight.detach()
        if conv.weight.is_contiguous(memory_format=torch.channels_last) and not conv.weight.is_contiguous():
            w = w.contiguous(memory_format=torch.channels_last)
        tw.weight = nn.Parameter(w, requires_grad=False)
        tw.bias = conv.bias                       # the same Parameter object
        tw.eval()
        self._twins[id(conv)] = (conv, tw)
        return tw

    def run(self, form: str, conv: nn.Conv2d, x: torch.Tensor, z=None, relu: bool = True) -> torch.Tensor:
        if form.endswith('+pad'):
            tw = self._twin(conv)
            x = F.pad(x, (0, 0, 0, 0, 0, tw.in_channels - x.shape[1]))        # keeps x's memory format
            conv, form = tw, form[:-4]
        if form == 'cudnn':
            return self.fused(conv, x, z)
        if form == 'kernel':
            return self.kernel(conv, x, z, relu)
        return self.unfused(conv, x, z, relu)

    # -- one-off trial per (layer, geometry, epilogue) -------------------------------------------------------
    @staticmethod
    def _key(conv, x, z, relu):
        return (id(conv), tuple(x.shape), tuple(x.stride()), x.dtype, z is not None, bool(relu),
                torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)

    def _time(self, fn) -> float:
        from cutie_b200.utils.op_trials import gpu_time_ms
        return gpu_time_ms(fn, self.trial_iters)       # launches queued back to back (device time, as in a graph replay)

    def _candidates(self, relu: bool, conv: nn.Conv2d = None):
        base = [f for f in self.forms if f != 'aten' and (relu or f != 'cudnn')]
        if conv is not None and self._padded_channels(conv):
            base += [f + '+pad' for f in self.forms if relu or f != 'cudnn']
        return base

    def _trial(self, key, conv, x, z, relu) -> str:
        from cutie_b200.kernels import KernelError
        what = f'conv {tuple(conv.weight.shape)} on {tuple(x.shape)}'
        ref = self.unfused(conv, x, z, relu)
        scale = float(ref.abs().max()) + 1e-6
        times = {'aten': self._time(lambda: self.unfused(conv, x, z, relu))}
        for form in self._candidates(relu, conv):
            try:
                out = self.run(form, conv, x, z, relu)
                err = float((out - ref).abs().max())
                # 'kernel' repeats ATen's arithmetic on the same cuDNN call; 'cudnn' may pick another engine (another
                # summation order / TF32 path) -- the check is against gross errors (layout, operand order), not rounding
                tol = (2e-2 if torch.backends.cudnn.allow_tf32 else 2e-4) * scale
                if not (err <= tol):               # also catches NaN
                    self.errors.append(f'{what}: {form} differs by {err:.3e} (scale {scale:.3e})')
                    continue
                times[form] = self._time(lambda: self.run(form, conv, x, z, relu))
            except KernelError:                    # our library missing / a failed launch is never absorbed
                raise
            except Exception as e:                 # noqa: BLE001 -- any cuDNN / dispatcher failure: drop the form
                self.errors.append(f'{what}: {form}: {type(e).__name__}: {e}')
        self.timings[key] = times
        return min(times, key=times.get)

    def _eligible(self, conv: nn.Conv2d, x: torch.Tensor) -> bool:
        return (self.enabled and x.is_cuda and conv.bias is not None and conv.padding_mode == 'zeros'
                and x.dim() == 4 and x.dtype == torch.float32 and not torch.is_grad_enabled())

    @staticmethod
    def _capturing() -> bool:
        return torch.cuda.is_current_stream_capturing()

    def __call__(self, conv: nn.Conv2d, x: torch.Tensor, z=None, relu: bool = True) -> torch.Tensor:
        if not self._eligible(conv, x):
            return self.unfused(conv, x, z, relu)
        key = self._key(conv, x, z, relu)
        form = self.decisions.get(key)
        if form is None:
            if self._capturing():
                return self.unfused(conv, x, z, relu)   # no timing inside a capture; _Captured warms up outside one first
            form = self.decisions[key] = self._trial(key, conv, x, z, relu)
        return self.run(form, conv, x, z, relu)

    # -- ResNet stem: relu(conv(x) + bias) -> max_pool2d(3, 2, 1) ---------------------------------------------
    def stem(self, conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
        """max_pool2d(relu(conv(x)), 3, stride=2, padding=1): either the chosen conv+ReLU form followed by ATen's pooling
        ('aten'), or the bias-less convolution followed by cutie_bias_relu_maxpool ('pool', 'pool+pad'): bias, clamp and
        pooling in one pass over the convolution output (they commute with max, so the result is the same)."""
        def aten():
            return F.max_pool2d(self(conv, x, None, True), 3, stride=2, padding=1)
        if not self._eligible(conv, x):
            return F.max_pool2d(self.unfused(conv, x), 3, stride=2, padding=1)
        key = ('stem',) + self._key(conv, x, None, True)
        form = self.decisions.get(key)
        if form is None:
            if self._capturing():
                return aten()
            form = self.decisions[key] = self._stem_trial(key, conv, x, aten)
        return self._stem_run(form, conv, x, aten)

    def _stem_run(self, form: str, conv: nn.Conv2d, x: torch.Tensor, aten) -> torch.Tensor:
        if form == 'aten':
            return aten()
        from cutie_b200 import kernels as K_
        if form.endswith('+pad'):
            tw = self._twin(conv)
            x = F.pad(x, (0, 0, 0, 0, 0, tw.in_channels - x.shape[1]))
            conv = tw
        return K_.bias_relu_maxpool(self._conv(conv, x, False), conv.bias)

    def _stem_trial(self, key, conv, x, aten) -> str:
        from cutie_b200.kernels import KernelError
        what = f'stem conv {tuple(conv.weight.shape)} on {tuple(x.shape)}'
        ref = aten()                                   # also settles the conv+ReLU decision it is built on
        scale = float(ref.abs().max()) + 1e-6
        times = {'aten': self._time(aten)}
        for form in ['pool'] + (['pool+pad'] if self._padded_channels(conv) else []):
            try:
                out = self._stem_run(form, conv, x, aten)
                err = float((out - ref).abs().max())
                tol = (2e-2 if torch.backends.cudnn.allow_tf32 else 2e-4) * scale
                if not (out.shape == ref.shape and err <= tol):
                    self.errors.append(f'{what}: {form} differs by {err:.3e} (scale {scale:.3e})')
                    continue
                times[form] = self._time(lambda: self._stem_run(form, conv, x, aten))
            except KernelError:
                raise
            except Exception as e:                     # noqa: BLE001 -- cuDNN / dispatcher failure of the bias-less call
                self.errors.append(f'{what}: {form}: {type(e).__name__}: {e}')
        self.timings[key] = times
        return min(times, key=times.get)

    def __deepcopy__(self, memo):          # a copied model gets its own (empty) fuser with the same settings
        new = ConvEpilogueFuser(self.enabled, self.trial_iters, self.forms)
        memo[id(self)] = new
        return new

    def report(self) -> dict:
        counts = {f: sum(1 for v in self.decisions.values() if v == f) for f in self.FORMS}
        counts['padded_input'] = sum(1 for v in self.decisions.values() if v.endswith('+pad'))
        counts['stem_pool'] = sum(1 for v in self.decisions.values() if v.startswith('pool'))
        saved = sum(t['aten'] - t[self.decisions[k]] for k, t in self.timings.items() if k in self.decisions)
        return {'enabled': self.enabled, **counts, 'errors': len(self.errors),
                'first_error': self.errors[0] if self.errors else None, 'trial_ms_saved_per_pass': saved}


def attach_epilogue_fuser(module: nn.Module, fuser: 'ConvEpilogueFuser') -> int:
    """Hands `fuser` to every nn.Conv2d under `module` (plain attributes, not parameters / buffers / sub-modules, so
    state_dict and .to() are unaffected).  Per model, not process-wide: an un-optimised model keeps PyTorch's launches.
    Besides `epilogue_fuser` (read by conv_relu / conv_add_relu) each convolution's `_conv_forward` is pointed at the
    fuser, so plain `conv(x)` calls -- no ReLU behind them -- get their bias from the fuser's choice as well."""
    n = 0
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            object.__setattr__(m, 'epilogue_fuser', fuser)
            object.__setattr__(m, '_conv_forward', _BiasOnlyForward(m, fuser))
            n += 1
    return n


class _BiasOnlyForward:
    """Instance-level replacement of nn.Conv2d._conv_forward(input, weight, bias) (called by Conv2d.forward)."""

    def __init__(self, conv: nn.Conv2d, fuser: ConvEpilogueFuser):
        self.conv, self.fuser = conv, fuser

    def __call__(self, x, weight, bias):
        conv = self.conv
        if bias is None or weight is not conv.weight or bias is not conv.bias:
            return nn.Conv2d._conv_forward(conv, x, weight, bias)
        return self.fuser(conv, x, None, relu=False)

    def __deepcopy__(self, memo):          # a copied module must point at ITS convolution (already in memo) and fuser
        import copy
        return _BiasOnlyForward(copy.deepcopy(self.conv, memo), copy.deepcopy(self.fuser, memo))


def conv_relu(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """relu(conv(x)) -- in the form the model's fuser chose for this layer, else convolution + bias add + clamp."""
    f = getattr(conv, 'epilogue_fuser', None)
    return ConvEpilogueFuser.unfused(conv, x) if f is None else f(conv, x)


def conv_relu_maxpool(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """max_pool2d(relu(conv(x)), 3, stride=2, padding=1) -- the ResNet stem."""
    f = getattr(conv, 'epilogue_fuser', None)
    if f is None:
        return F.max_pool2d(ConvEpilogueFuser.unfused(conv, x), 3, stride=2, padding=1)
    return f.stem(conv, x)


def conv_add(conv: nn.Conv2d, x: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """conv(x) + z (bias included, no activation): the closing convolution of a pre-activation residual block."""
    f = getattr(conv, 'epilogue_fuser', None)
    return ConvEpilogueFuser.unfused(conv, x, z, relu=False) if f is None else f(conv, x, z, relu=False)


def conv_plain(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """conv(x) on a 4-D tensor for any nn.Conv2d subclass (ObjConv2d's own forward expects 5-D object tensors)."""
    f = getattr(conv, 'epilogue_fuser', None)
    return ConvEpilogueFuser._conv(conv, x, True) if f is None else f(conv, x, None, relu=False)


def conv_add_relu(conv: nn.Conv2d, x: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """relu(conv(x) + z)."""
    f = getattr(conv, 'epilogue_fuser', None)
    return ConvEpilogueFuser.unfused(conv, x, z) if f is None else f(conv, x, z)
