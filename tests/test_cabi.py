"""The C-ABI library builds for sm_100a here (nvcc cross-compiles without a GPU), loads, and exports
every symbol include/cutie_b200.h declares.  No compute calls (no GPU in this suite)."""
import ctypes
import os
import re
import subprocess

from tests.conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, 'include', 'cutie_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(cutie_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_loads_and_exports_header_symbols():
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(ge.LIB)
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/cutie_b200.h but not exported'
    assert lib.cutie_b200_abi_version() == 1
    lib.cutie_b200_last_error.restype = ctypes.c_char_p
    assert lib.cutie_b200_last_error() is not None


def test_argument_validation_needs_no_gpu():
    """Invalid arguments are rejected before any CUDA call and set the thread-local error string."""
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(ge.LIB)
    lib.cutie_b200_last_error.restype = ctypes.c_char_p
    st = lib.cutie_obj_summary_accumulate(None, None, ctypes.c_int64(4), None)
    assert st == -1 and b'cutie_obj_summary_accumulate' in lib.cutie_b200_last_error()
    st = lib.cutie_qt_self_attention(None, None, ctypes.c_int64(16), ctypes.c_int64(256), 16, 8, None, None)
    assert st == -1


def test_sass_is_sm100a():
    import __graft_entry__ as ge
    ge.build()
    out = subprocess.run(['/usr/local/cuda/bin/cuobjdump', '-lelf', ge.LIB], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_affinity_plan_is_pure_host_logic():
    """Which passes cutie_affinity_topk runs for a bank size (no GPU needed): exact scan below the threshold,
    nested tcgen05 filter levels (strides 16^l) above it, coarsest sample never above 4096 tokens."""
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(ge.LIB)
    plan = lambda n, k=30: lib.cutie_affinity_plan_levels(ctypes.c_int64(n), k)
    lib.cutie_set_tc_min_tokens(ctypes.c_int64(-1))
    assert plan(100) == 0 and plan(1620) == 0 and plan(4860) == 0
    assert plan(8100) == 2            # 8100/16 = 507-token all-pass sample, then the whole bank
    assert plan(65536) == 2           # 4096-token sample
    assert plan(65537) == 3 and plan(413100) == 3 and plan(414720) == 3
    assert plan(20_000_000) == 5       # strides 65536, 4096, 256, 16, 1
    lib.cutie_set_tc_min_tokens(ctypes.c_int64(256))
    assert plan(333) == 1 and plan(59) == 0 and plan(4099) == 2
    lib.cutie_set_tc_min_tokens(ctypes.c_int64(1 << 40))
    assert plan(413100) == 0
    lib.cutie_set_tc_min_tokens(ctypes.c_int64(-1))
