"""First-use A/B of an ATen composition against one of our kernels (per op and geometry).

The pixel-side glue kernels in csrc/pixel.cu replace short chains of ATen launches inside the PyTorch stages around
the hot path.  Which form is faster depends on the GPU, the driver and the shapes, so nothing is assumed: the first
time an (op, geometry) pair is seen OUTSIDE a stream capture both forms run on the live tensors, the kernel's result
must match the ATen composition, both are timed with CUDA events, and the faster one is kept (`decisions`).  A
mismatch keeps the ATen form and is recorded in `errors` (bench.py prints the report).  A missing library or a failed
launch (`KernelError`) is NOT absorbed -- it propagates like everywhere else in cutie_b200.  CPU tensors (the oracle
harness borrowing the modules) always take the ATen form.

Attached per model by `CUTIE.optimize_for_inference()` (attribute `op_trials` on every sub-module); modules without
it run PyTorch's launches unchanged.
"""
from typing import Callable, Dict

import torch
import torch.nn as nn


def gpu_time_ms(fn: Callable, iters: int) -> float:
    """Device time of one fn() with its launches queued back to back.  Measured eagerly, a chain of small kernels is
    bounded by the host's launch rate (the GPU idles between them), which would favour whichever form has fewer
    launches even when its kernels are slower -- but the frame path replays these launches from CUDA graphs, where only
    device time counts.  So the GPU is parked on a ~1.5 ms spin kernel first, the host runs ahead and queues all
    iterations behind it, and the events bracket a gap-free execution."""
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        torch.cuda._sleep(3_000_000)
    except Exception:              # noqa: BLE001 -- private helper missing: plain eager timing
        pass
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


class OpTrials:
    def __init__(self, enabled: bool = True, trial_iters: int = 6):
        self.enabled = enabled
        self.trial_iters = trial_iters
        self.decisions: Dict[tuple, bool] = {}       # (op, key) -> True: our kernel
        self.timings: Dict[tuple, tuple] = {}        # (op, key) -> (kernel_ms, aten_ms)
        self.errors = []

    def _eligible(self, probe: torch.Tensor) -> bool:
        return self.enabled and probe.is_cuda and probe.dtype == torch.float32 and not torch.is_grad_enabled()

    @staticmethod
    def _capturing() -> bool:
        return torch.cuda.is_current_stream_capturing()

    def _time(self, fn) -> float:
        return gpu_time_ms(fn, self.trial_iters)

    def _trial(self, op, key, aten: Callable, kernel: Callable, rtol: float) -> bool:
        from cutie_b200.kernels import KernelError
        ref = aten()
        try:
            out = kernel(True)                       # trial=True: in-place kernels work on a copy
        except KernelError:                          # our library missing / a failed launch is never absorbed
            raise
        except Exception as e:                       # noqa: BLE001 -- a PyTorch / cuDNN variant that does not run here
            self.errors.append(f'{op} {key}: {type(e).__name__}: {e}')
            return False
        refs = list(ref) if isinstance(ref, (tuple, list)) else [ref]
        outs = list(out) if isinstance(out, (tuple, list)) else [out]
        ok = len(refs) == len(outs)
        worst = 0.0
        for r, o in zip(refs, outs):
            scale = float(r.abs().max()) + 1e-6
            err = float((o - r).abs().max()) if o.shape == r.shape else float('nan')
            worst = max(worst, err / scale) if err == err else float('nan')
            ok = ok and o.shape == r.shape and err <= rtol * scale            # also catches NaN
        if not ok:
            self.errors.append(f'{op} {key}: kernel differs from ATen (relative error {worst:.3e}, allowed {rtol:.1e})')
            return False
        t_k = self._time(lambda: kernel(True))
        t_a = self._time(aten)
        self.timings[(op, key)] = (t_k, t_a)
        return t_k <= t_a

    def __call__(self, op: str, key: tuple, aten: Callable, kernel: Callable, probe: torch.Tensor,
                 rtol: float = 1e-5):
        """aten(): the PyTorch composition (a tensor or a tuple of tensors).  kernel(trial: bool): our kernel, same
        outputs; with trial=True it must not modify its inputs (in-place kernels clone their destination)."""
        if not self._eligible(probe):
            return aten()
        use = self.decisions.get((op, key))
        if use is None:
            if self._capturing():
                return aten()
            use = self.decisions[(op, key)] = self._trial(op, key, aten, kernel, rtol)
        return kernel(False) if use else aten()

    def pick(self, op: str, key: tuple, candidates, run: Callable, probe: torch.Tensor):
        """Launch-parameter A/B: `run(c)` performs the op with parameter c (all candidates are valid; results may differ
        in summation order only).  Returns the fastest candidate for (op, key); candidates[0] -- the library default --
        for CPU tensors, inside a capture before a decision exists, or when there is nothing to choose."""
        if len(candidates) < 2 or not self._eligible(probe):
            return candidates[0]
        got = self.decisions.get((op, key))
        if got is None:
            if self._capturing():
                return candidates[0]
            times = [self._time(lambda c=c: run(c)) for c in candidates]
            best = min(range(len(candidates)), key=times.__getitem__)
            got = self.decisions[(op, key)] = candidates[best]
            self.timings[(op, key)] = (times[best], times[0])
        return got

    def __deepcopy__(self, memo):
        new = OpTrials(self.enabled, self.trial_iters)
        memo[id(self)] = new
        return new

    def report(self) -> dict:
        ops = {}
        for (op, _), use in self.decisions.items():
            if isinstance(use, bool):
                d = ops.setdefault(op, {'kernel': 0, 'aten': 0})
                d['kernel' if use else 'aten'] += 1
            else:                                             # a picked launch parameter
                ops.setdefault(op, {'picked': []})['picked'].append(use)
        saved = sum(a - k for key, (k, a) in self.timings.items() if self.decisions.get(key) not in (None, False))
        return {'enabled': self.enabled, 'ops': ops, 'errors': len(self.errors),
                'first_error': self.errors[0] if self.errors else None, 'trial_ms_saved_per_pass': saved}


def attach_op_trials(module: nn.Module, trials: OpTrials) -> int:
    n = 0
    for m in module.modules():
        object.__setattr__(m, 'op_trials', trials)
        n += 1
    return n
