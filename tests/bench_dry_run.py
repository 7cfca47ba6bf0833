"""Test-only driver: runs bench.py's control flow on a machine without a GPU.

torch.cuda's stream / event / graph entry points are replaced by inert stand-ins, bench's `torch.device('cuda', i)`
resolves to the CPU, the kernel wrappers are the CPU emulations of tests/cpu_kernels.py and the workload is a tiny clip.
Nothing measured here means anything; the point is that every line of bench.py (arms, extension arm, parity check, JSON
assembly) executes before it first runs on a B200.  Usage: python tests/bench_dry_run.py [bench args]"""
import contextlib
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)

    def synchronize(self):
        pass

    def wait(self, stream=None):
        pass


class _Stream:
    def __init__(self, device=None, **kw):
        self.device = device if isinstance(device, torch.device) else torch.device('cpu')
        self.cuda_stream = 0

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass

    def synchronize(self):
        pass


_cur = _Stream()
torch.cuda.Event = _Event
torch.cuda.Stream = _Stream
torch.cuda.current_stream = lambda device=None: _cur
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.cuda.synchronize = lambda device=None: None
torch.cuda.set_device = lambda d: None
torch.cuda.is_available = lambda: True
torch.cuda.memory_allocated = lambda device=None: 0
torch.cuda.is_current_stream_capturing = lambda: False
torch.cuda._sleep = lambda n: None
torch.Tensor.pin_memory = lambda self, *a, **k: self
torch.Tensor.record_stream = lambda self, s: None

from tests import cpu_kernels  # noqa: E402
cpu_kernels.install()

import cutie_b200.kernels as K_  # noqa: E402


def _timed(name, fn, launches):
    def wrapped(*a, **k):
        with K_._call(name, launches):          # so that bench's per-kernel table and roofline assembly run too
            return fn(*a, **k)
    return wrapped


K_.affinity_topk = _timed('affinity_topk', K_.affinity_topk, 2)
K_.readout_gather = _timed('readout_gather', K_.readout_gather, 1)

# the CUDA-graph frame path and the encoder look-ahead with stand-in graphs / streams (tests/test_graph_path_cpu.py)
import cutie_b200.inference.frame_graphs as _fg  # noqa: E402
import cutie_b200.inference.inference_core as _ic  # noqa: E402
from tests.test_graph_path_cpu import _FakeCaptured, _NoStreams  # noqa: E402
_fg._Captured = _FakeCaptured
_ic._graphable = lambda t: True
_ic._CudaStreamOps = _NoStreams

sys.argv = ['bench.py'] + sys.argv[1:]
import bench  # noqa: E402  (points fd 1 at stderr; the JSON goes to the real stdout)


class _TorchProxy(types.ModuleType):
    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(*a, **k):
        return torch.device('cpu')


bench.torch = _TorchProxy('torch')
bench.WORKLOADS['tiny'] = dict(H=96, W=160, K=2, mem_frames=4, top_k=30, desc='dry run (CPU, emulated kernels)')
if '--workload' not in sys.argv:
    sys.argv += ['--workload', 'tiny']
if '--no-northstar' not in sys.argv:
    sys.argv += ['--no-northstar']          # a 10 000-key read through the CPU emulation would take minutes
bench.main()
