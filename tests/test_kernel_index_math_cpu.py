"""The index arithmetic of the barrier-free pixel-side kernels (cutie_bias_act, cutie_area_pool, cutie_eca_scale_add,
cutie_gated_update), executed on the host: each kernel TEMPLATE is cut out of csrc/pixel.cu verbatim, compiled by g++
behind a serial (block, thread) loop (tests/emul/cuda_serial_shim.h) and compared with PyTorch for NCHW /
channels-last storage, vector / scalar paths and sizes whose grid-stride loops wrap.  (No GPU: this checks
addressing and arithmetic, not performance.)"""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from tests.conftest import ROOT

HARNESS = r'''
#include "cuda_serial_shim.h"
namespace cutie {
%(kernel)s
%(more)s
}
extern "C" void emu_area_pool(const float* in, float* out, long long planes, int H, int W, int f, int variant) {
  const int Ho = H / f, Wo = W / f;
  const long long total = planes * Ho * Wo;
  const int blocks = (int)((total + 127) / 128);
  if (variant == 16) EMU_LAUNCH((cutie::area_pool_kernel<16, true>), blocks, 128, in, out, total, Ho, Wo, f);
  else if (variant == 4) EMU_LAUNCH((cutie::area_pool_kernel<4, true>), blocks, 128, in, out, total, Ho, Wo, f);
  else if (variant == 2) EMU_LAUNCH((cutie::area_pool_kernel<2, false>), blocks, 128, in, out, total, Ho, Wo, f);
  else EMU_LAUNCH((cutie::area_pool_kernel<0, false>), blocks, 128, in, out, total, Ho, Wo, f);
}
extern "C" void emu_eca(float* y, const float* x, const float* mean, const float* w, float* gate, long long N, long long C,
                        long long HW, int k, int cl, int vec, int blocks) {
  const long long nc = N * C;
  EMU_LAUNCH(cutie::eca_gate_kernel, (int)((nc + 255) / 256), 256, mean, w, gate, nc, (int)C, k);
  const long long n = nc * HW, total = vec ? n / 4 : n;
  if (cl) {
    if (vec) EMU_LAUNCH((cutie::scale_add_kernel<true, true>), blocks, 256, y, gate, x, total, (int)C, HW);
    else EMU_LAUNCH((cutie::scale_add_kernel<true, false>), blocks, 256, y, gate, x, total, (int)C, HW);
  } else {
    if (vec) EMU_LAUNCH((cutie::scale_add_kernel<false, true>), blocks, 256, y, gate, x, total, (int)C, HW);
    else EMU_LAUNCH((cutie::scale_add_kernel<false, false>), blocks, 256, y, gate, x, total, (int)C, HW);
  }
}
extern "C" void emu_stem_pool(const float* y, const float* bias, float* out, long long N, int C, int H, int W, int cl) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = cl ? N * Ho * Wo * (C / 4) : N * C * Ho * Wo;
  const int blocks = (int)((total + 255) / 256);
  if (cl) EMU_LAUNCH((cutie::bias_relu_maxpool_kernel<true>), blocks, 256, y, bias, out, total, C, H, W, Ho, Wo);
  else EMU_LAUNCH((cutie::bias_relu_maxpool_kernel<false>), blocks, 256, y, bias, out, total, C, H, W, Ho, Wo);
}
extern "C" void emu_segment_tail(const float* x, float* agg, float* logits, float* prob, long long B, int K, int h, int w) {
  const long long lo = B * h * w;
  EMU_LAUNCH(cutie::aggregate_logits_kernel, (int)((lo + 255) / 256), 256, x, agg, lo, K, (long long)h * w);
  const long long hi = lo * 16;
  EMU_LAUNCH(cutie::upsample4_softmax_kernel, (int)((hi + 255) / 256), 256, agg, logits, prob, hi, K + 1, h, w);
}
extern "C" void emu_gated(const float* v, const float* h, float* out, long long P, long long d, long long HW, int vec,
                          int blocks) {
  const long long n = P * d * HW, total = vec ? n / 4 : n;
  if (vec) EMU_LAUNCH((cutie::gated_update_kernel<true>), blocks, 256, v, h, out, total, (int)d, HW);
  else EMU_LAUNCH((cutie::gated_update_kernel<false>), blocks, 256, v, h, out, total, (int)d, HW);
}
extern "C" void emu_bias_act(float* y, const float* bias, const float* z, long long N, long long C, long long HW,
                             int cl, int vec, int relu, int blocks) {
  const long long n = N * C * HW;
  const long long total = vec ? n / 4 : n;
  if (cl) {
    if (vec) EMU_LAUNCH((cutie::bias_act_kernel<true, true>), blocks, 256, y, bias, z, total, (int)C, HW, relu);
    else EMU_LAUNCH((cutie::bias_act_kernel<true, false>), blocks, 256, y, bias, z, total, (int)C, HW, relu);
  } else {
    if (vec) EMU_LAUNCH((cutie::bias_act_kernel<false, true>), blocks, 256, y, bias, z, total, (int)C, HW, relu);
    else EMU_LAUNCH((cutie::bias_act_kernel<false, false>), blocks, 256, y, bias, z, total, (int)C, HW, relu);
  }
}
'''


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
    src = open(os.path.join(ROOT, 'cutie_b200', 'csrc', 'pixel.cu')).read()
    m = re.search(r'(template <bool CL, bool VEC>\n__global__ void .*?bias_act_kernel\(.*?\n}\n)', src, re.S)
    assert m, 'bias_act_kernel not found in pixel.cu'
    more = []
    for head in (r'template <int F, bool VEC>\n__global__ void [^\n]*area_pool_kernel\(',
                 r'__global__ void [^\n]*eca_gate_kernel\(',
                 r'template <bool CL, bool VEC>\n__global__ void [^\n]*scale_add_kernel\(',
                 r'template <bool VEC>\n__global__ void [^\n]*gated_update_kernel\(',
                 r'template <bool CL>\n__global__ void [^\n]*bias_relu_maxpool_kernel\(',
                 r'constexpr int SEG_MAXC = 16;\n\n__global__ void [^\n]*aggregate_logits_kernel\(',
                 r'__device__ __forceinline__ void src_index4\(',
                 r'__global__ void [^\n]*upsample4_softmax_kernel\('):
        mm = re.search('(' + head + r'.*?\n}\n)', src, re.S)
        assert mm, head
        more.append(mm.group(1))
    d = tmp_path_factory.mktemp('emu')
    cpp = d / 'emu_bias_act.cpp'
    cpp.write_text(HARNESS % {'kernel': m.group(1), 'more': '\n'.join(more)})
    so = d / 'libemu_bias_act.so'
    subprocess.run(['g++', '-O1', '-ffp-contract=off', '-shared', '-fPIC', '-std=c++17', '-I', os.path.join(ROOT, 'tests', 'emul'),
                    '-o', str(so), str(cpp)], check=True)
    lib = ctypes.CDLL(str(so))
    return lib


@pytest.mark.parametrize('shape,cl,vec', [
    ((2, 8, 6, 4), False, True), ((2, 8, 6, 4), False, False), ((2, 8, 6, 4), True, True), ((2, 8, 6, 4), True, False),
    ((3, 5, 3, 3), False, False), ((3, 5, 3, 3), True, False), ((1, 1, 30, 54), False, True),
    ((2, 12, 5, 4), True, True), ((2, 3, 9, 4), False, True),
])
@pytest.mark.parametrize('relu', [0, 1])
@pytest.mark.parametrize('with_z', [False, True])
@pytest.mark.parametrize('blocks', [1, 3])          # 1 block of 256 threads: the grid-stride loop wraps at these sizes
def test_bias_act_index_math(emu, shape, cl, vec, relu, with_z, blocks):
    N, C, H, W = shape
    g = torch.Generator().manual_seed(N * 1000 + C * 100 + H * 10 + W + relu)
    y = torch.randn(*shape, generator=g)
    z = torch.randn(*shape, generator=g) if with_z else None
    bias = torch.randn(C, generator=g)
    want = y + bias.view(1, -1, 1, 1)
    if z is not None:
        want = want + z
    if relu:
        want = torch.relu(want)
    if cl:      # storage [N, HW, C]
        ybuf = y.permute(0, 2, 3, 1).contiguous()
        zbuf = z.permute(0, 2, 3, 1).contiguous() if with_z else None
    else:
        ybuf = y.contiguous().clone()
        zbuf = z.contiguous() if with_z else None
    P = ctypes.c_void_p
    emu.emu_bias_act(P(ybuf.data_ptr()), P(bias.data_ptr()), P(zbuf.data_ptr() if with_z else 0),
                     ctypes.c_longlong(N), ctypes.c_longlong(C), ctypes.c_longlong(H * W), int(cl), int(vec), relu, blocks)
    got = ybuf.permute(0, 3, 1, 2) if cl else ybuf
    assert torch.equal(got, want)


@pytest.mark.parametrize('lead,H,W,f,variant', [
    ((3,), 32, 48, 16, 16), ((2, 3), 8, 12, 4, 4), ((2, 5), 6, 10, 2, 2), ((4,), 9, 6, 3, 0), ((1,), 32, 48, 16, 0),
    ((2,), 8, 12, 4, 0), ((300,), 4, 4, 2, 2),
])
def test_area_pool_index_math(emu, lead, H, W, f, variant):
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(H * W + f)
    x = torch.randn(*lead, H, W, generator=g)
    planes = x.numel() // (H * W)
    out = torch.full((*lead, H // f, W // f), float('nan'))
    P = ctypes.c_void_p
    emu.emu_area_pool(P(x.data_ptr()), P(out.data_ptr()), ctypes.c_longlong(planes), H, W, f, variant)
    want = F.interpolate(x.reshape(-1, 1, H, W), size=(H // f, W // f), mode='area').reshape(out.shape)
    assert torch.allclose(out, want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('shape,cl,vec', [((2, 8, 3, 4), False, True), ((2, 8, 3, 4), False, False), ((2, 8, 3, 4), True, True),
                                          ((2, 8, 3, 4), True, False), ((3, 6, 5, 3), True, False), ((3, 6, 5, 3), False, False),
                                          ((2, 260, 2, 2), False, True), ((2, 260, 2, 2), True, True)])
@pytest.mark.parametrize('k', [1, 3, 5])
@pytest.mark.parametrize('blocks', [1, 2])
def test_eca_scale_add_index_math(emu, shape, cl, vec, k, blocks):
    import torch.nn.functional as F
    N, C, H, W = shape
    g = torch.Generator().manual_seed(C * 7 + k)
    y, x = torch.randn(*shape, generator=g), torch.randn(*shape, generator=g)
    w = torch.randn(1, 1, k, generator=g)
    mean = y.mean(dim=(2, 3)).contiguous()
    gate_want = torch.sigmoid(F.conv1d(mean.unsqueeze(1), w, padding=(k - 1) // 2)).squeeze(1)
    want = y * gate_want.view(N, C, 1, 1) + x
    if cl:
        ybuf, xbuf = y.permute(0, 2, 3, 1).contiguous(), x.permute(0, 2, 3, 1).contiguous()
    else:
        ybuf, xbuf = y.clone(), x.clone()
    gate = torch.full((N, C), float('nan'))
    P = ctypes.c_void_p
    emu.emu_eca(P(ybuf.data_ptr()), P(xbuf.data_ptr()), P(mean.data_ptr()), P(w.data_ptr()), P(gate.data_ptr()),
                ctypes.c_longlong(N), ctypes.c_longlong(C), ctypes.c_longlong(H * W), k, int(cl), int(vec), blocks)
    got = ybuf.permute(0, 3, 1, 2) if cl else ybuf
    assert torch.allclose(gate, gate_want, rtol=1e-6, atol=1e-7)
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('P_,d,HW,vec', [(3, 4, 8, True), (3, 4, 8, False), (2, 5, 9, False), (6, 16, 20, True)])
@pytest.mark.parametrize('blocks', [1, 3])
def test_gated_update_index_math(emu, P_, d, HW, vec, blocks):
    g = torch.Generator().manual_seed(P_ * 100 + d * 10 + HW)
    v = torch.randn(P_, 3 * d, HW, generator=g) * 2
    h = torch.randn(P_, d, HW, generator=g)
    f, u, n = torch.sigmoid(v[:, :d]), torch.sigmoid(v[:, d:2 * d]), torch.tanh(v[:, 2 * d:])
    want = f * h * (1 - u) + u * n
    out = torch.full_like(h, float('nan'))
    P = ctypes.c_void_p
    emu.emu_gated(P(v.data_ptr()), P(h.data_ptr()), P(out.data_ptr()), ctypes.c_longlong(P_), ctypes.c_longlong(d),
                  ctypes.c_longlong(HW), int(vec), blocks)
    assert torch.allclose(out, want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('shape', [(2, 8, 9, 12), (1, 4, 6, 5), (3, 12, 7, 7), (1, 64, 30, 54), (2, 4, 1, 1), (1, 8, 2, 3)])
@pytest.mark.parametrize('cl', [False, True])
def test_bias_relu_maxpool_index_math(emu, shape, cl):
    import torch.nn.functional as F
    N, C, H, W = shape
    g = torch.Generator().manual_seed(H * 31 + W)
    y = torch.randn(*shape, generator=g)
    bias = torch.randn(C, generator=g)
    want = F.max_pool2d(torch.relu(y + bias.view(1, -1, 1, 1)), 3, stride=2, padding=1)
    Ho, Wo = want.shape[-2:]
    assert (Ho, Wo) == ((H - 1) // 2 + 1, (W - 1) // 2 + 1)
    P = ctypes.c_void_p
    if cl:
        ybuf = y.permute(0, 2, 3, 1).contiguous()
        out = torch.full((N, Ho, Wo, C), float('nan'))
    else:
        ybuf = y.contiguous()
        out = torch.full((N, C, Ho, Wo), float('nan'))
    emu.emu_stem_pool(P(ybuf.data_ptr()), P(bias.data_ptr()), P(out.data_ptr()), ctypes.c_longlong(N), C, H, W, int(cl))
    got = out.permute(0, 3, 1, 2) if cl else out
    assert torch.equal(got, want)               # max, one add, clamp: bit-identical to maxpool(relu(y + b))


@pytest.mark.parametrize('B,K,h,w', [(1, 3, 6, 9), (2, 1, 3, 4), (1, 15, 2, 2), (1, 5, 1, 7)])
def test_segment_tail_index_math(emu, B, K, h, w):
    import torch.nn.functional as F
    from cutie_b200.utils.tensor_utils import aggregate
    g = torch.Generator().manual_seed(K * 10 + h)
    x = 4 * torch.randn(B, K, h, w, generator=g)
    x[0, 0, 0, 0] = 40.0                                   # saturates the clamp at 1 - 1e-7
    x[0, K - 1, -1, -1] = -40.0                            # and at 1e-7
    agg_want = aggregate(torch.sigmoid(x), dim=1)
    lg_want = F.interpolate(agg_want, scale_factor=4, mode='bilinear', align_corners=False)
    pr_want = F.softmax(lg_want, dim=1)
    agg = torch.full((B, K + 1, h, w), float('nan'))
    lg = torch.full((B, K + 1, 4 * h, 4 * w), float('nan'))
    pr = torch.full_like(lg, float('nan'))
    P = ctypes.c_void_p
    emu.emu_segment_tail(P(x.data_ptr()), P(agg.data_ptr()), P(lg.data_ptr()), P(pr.data_ptr()), ctypes.c_longlong(B), K, h, w)
    assert torch.allclose(agg, agg_want, rtol=1e-5, atol=1e-5)
    assert torch.allclose(lg, lg_want, rtol=1e-5, atol=1e-5)
    assert torch.allclose(pr, pr_want, rtol=1e-5, atol=1e-6)
