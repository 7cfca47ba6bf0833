"""Test-only: compile csrc/pixel.cu for the HOST.  The kernels in that file are barrier-free or meet at a single
__syncthreads() after an idempotent first half (conv3x3_c1's partial sums), so `kernel<<<grid, block, smem, stream>>>(args)`
can be rewritten textually into a serial loop over (block, thread) -- two passes per block for the one-barrier kernels
(cuda_serial_shim.h) -- and the whole translation unit -- kernels AND the extern "C" dispatchers (vector / scalar path
selection, grid sizing, argument checks) -- builds with g++.  The resulting library exports the same C-ABI symbols as
libcutie_b200.so for those entry points, so the real ctypes wrappers in cutie_b200/kernels.py can be driven on CPU
tensors (tests/test_pixel_wrappers_on_host.py)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

PRELUDE = r'''
#include "cuda_serial_shim.h"
#include <stdint.h>
typedef void* cudaStream_t;
#define CUTIE_REQUIRE(cond, what) do { if (!(cond)) return -1; } while (0)
#define CUTIE_CHECK_LAUNCH() do { } while (0)
namespace cutie { static inline int num_sms() { return 148; } }
'''


def host_source() -> str:
    src = open(os.path.join(ROOT, 'cutie_b200', 'csrc', 'pixel.cu')).read()
    src = src.replace('#include <math_constants.h>', '').replace('#include "common.cuh"', PRELUDE)
    launch = re.compile(r'(\b[A-Za-z_][\w:]*(?:<[^<>;()]*>)?)\s*<<<\s*([^;]*?)>>>\s*\(', re.S)

    def rewrite(m):
        cfg = [c.strip() for c in split_top_level(m.group(2))]
        return f'EMU_LAUNCH(({m.group(1)}), {cfg[0]}, {cfg[1]}, '
    out, n = launch.subn(rewrite, src)
    assert n >= 8, f'only {n} kernel launches rewritten'
    return out


def split_top_level(s: str):
    parts, depth, cur = [], 0, ''
    for ch in s:
        if ch in '([<':
            depth += 1
        elif ch in ')]>':
            depth -= 1
        if ch == ',' and depth == 0:
            parts.append(cur)
            cur = ''
        else:
            cur += ch
    parts.append(cur)
    return parts


def build(out_dir: str) -> str:
    cpp = os.path.join(out_dir, 'pixel_host.cpp')
    so = os.path.join(out_dir, 'libpixel_host.so')
    with open(cpp, 'w') as f:
        f.write(host_source())
    subprocess.run(['g++', '-O1', '-ffp-contract=off', '-shared', '-fPIC', '-std=c++17', '-I',
                    os.path.join(ROOT, 'tests', 'emul'), '-o', so, cpp], check=True)
    return so
