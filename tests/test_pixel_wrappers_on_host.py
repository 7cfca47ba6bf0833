"""The real ctypes wrappers of cutie_b200/kernels.py for the pixel-side kernels, driven on CPU tensors against a HOST
build of csrc/pixel.cu (tests/emul/host_build.py: every `<<<...>>>` launch rewritten into a serial loop, kernels and
extern "C" dispatchers compiled verbatim by g++).  What runs here is the code that runs on the GPU minus the hardware:
Python argument handling (layouts, copies, shapes), the C dispatch (vector / scalar selection on alignment and sizes,
grid computation, argument checks) and the kernel bodies' index arithmetic -- compared with the PyTorch ops they
replace.  (No GPU; performance and memory-model behaviour are not what this checks.)"""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from tests.emul import host_build


@pytest.fixture(scope='module')
def hostlib(tmp_path_factory):
    return ctypes.CDLL(host_build.build(str(tmp_path_factory.mktemp('pixel_host'))))


@pytest.fixture
def K_(hostlib, monkeypatch):
    import cutie_b200.kernels as k

    def ptr(t, dtype=torch.float32):
        if t is None:
            return ctypes.c_void_p(0)
        assert t.dtype == dtype, (t.dtype, dtype)
        return ctypes.c_void_p(t.data_ptr())
    monkeypatch.setattr(k, 'lib', lambda: hostlib)
    monkeypatch.setattr(k, '_ptr', ptr)
    monkeypatch.setattr(k, '_stream', lambda: ctypes.c_void_p(0))
    return k


def _fmt(cl):
    return torch.channels_last if cl else torch.contiguous_format


@pytest.mark.parametrize('shape', [(2, 8, 6, 4), (3, 5, 3, 3), (1, 1, 30, 54), (2, 12, 5, 4), (1, 64, 9, 11)])
@pytest.mark.parametrize('cl', [False, True])
@pytest.mark.parametrize('relu', [False, True])
@pytest.mark.parametrize('with_z', [False, True])
def test_bias_act_wrapper(K_, shape, cl, relu, with_z):
    g = torch.Generator().manual_seed(sum(shape))
    y = torch.randn(*shape, generator=g).contiguous(memory_format=_fmt(cl))
    z = torch.randn(*shape, generator=g) if with_z else None          # always NCHW: the wrapper re-lays it out when cl
    b = torch.randn(shape[1], generator=g)
    want = y + b.view(1, -1, 1, 1)
    want = want + z if with_z else want
    want = torch.relu(want) if relu else want
    got = K_.bias_act_(y.clone(memory_format=torch.preserve_format), b, z, relu)
    assert got.stride() == y.stride() and torch.equal(got, want)


def test_bias_act_wrapper_misaligned_view_and_errors(K_):
    base = torch.randn(1 + 2 * 8 * 6 * 4)
    y = base[1:].view(2, 8, 6, 4)                                     # 4-byte aligned only -> scalar kernel
    b = torch.randn(8)
    want = torch.relu(y + b.view(1, -1, 1, 1))
    assert torch.equal(K_.bias_act_(y, b, None, True), want)
    with pytest.raises(K_.KernelError):
        K_.bias_act_(torch.randn(2, 8, 6, 4)[:, :, ::2], b)           # not dense


@pytest.mark.parametrize('shape,f', [((1, 3, 32, 48), 16), ((3, 16, 8, 12), 2), ((2, 5, 8, 12), 4), ((2, 5, 9, 12), 3),
                                     ((3, 32, 48), 16), ((2, 3, 6, 10), 2)])
def test_area_pool_wrapper(K_, shape, f):
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(f))
    H, W = shape[-2:]
    want = F.interpolate(x.reshape(-1, 1, H, W), size=(H // f, W // f), mode='area').reshape(*shape[:-2], H // f, W // f)
    got = K_.area_pool(x, f)
    assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-6, atol=1e-6)
    xt = x.transpose(-1, -2).contiguous().transpose(-1, -2)            # non-contiguous input: the wrapper copies
    assert torch.allclose(K_.area_pool(xt, f), want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('shape', [(3, 16, 6, 9), (2, 8, 5, 3), (1, 260, 2, 2)])
@pytest.mark.parametrize('cl', [False, True])
def test_eca_scale_add_wrapper(K_, shape, cl):
    g = torch.Generator().manual_seed(shape[1])
    y = torch.randn(*shape, generator=g).contiguous(memory_format=_fmt(cl))
    x = torch.randn(*shape, generator=g)
    conv = torch.nn.Conv1d(1, 1, 5, padding=2, bias=False)
    with torch.no_grad():
        gate = conv(y.mean(dim=(2, 3)).unsqueeze(1)).sigmoid().transpose(1, 2).unsqueeze(-1)
        want = y * gate + x
        got = K_.eca_scale_add_(y.clone(memory_format=torch.preserve_format), x, conv.weight)
    assert got.stride() == y.stride() and torch.allclose(got, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('shape', [(1, 3, 16, 6, 9), (2, 2, 4, 5, 3)])
def test_gated_update_wrapper(K_, shape):
    from cutie_b200.model.blocks import gated_update
    B, K, d, H, W = shape
    g = torch.Generator().manual_seed(d)
    h = torch.randn(*shape, generator=g)
    v = 2 * torch.randn(B, K, 3 * d, H, W, generator=g)
    want = gated_update(h, v)
    assert torch.allclose(K_.gated_update(h, v), want, rtol=1e-6, atol=1e-6)
    vt = v.transpose(-1, -2).contiguous().transpose(-1, -2)            # strided v: the wrapper copies
    assert torch.allclose(K_.gated_update(h, vt), want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('shape', [(1, 64, 12, 20), (3, 8, 9, 7), (1, 6, 7, 7), (2, 4, 1, 1)])
@pytest.mark.parametrize('cl', [False, True])
def test_bias_relu_maxpool_wrapper(K_, shape, cl):
    g = torch.Generator().manual_seed(shape[2])
    y = torch.randn(*shape, generator=g).contiguous(memory_format=_fmt(cl))
    b = torch.randn(shape[1], generator=g)
    want = F.max_pool2d(torch.relu(y + b.view(1, -1, 1, 1)), 3, stride=2, padding=1)
    got = K_.bias_relu_maxpool(y, b)
    assert got.shape == want.shape and torch.equal(got, want)
    if cl and shape[1] % 4 == 0 and shape[1] > 1 and shape[2] * shape[3] > 1:
        assert got.is_contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize('B,K,h,w', [(1, 3, 6, 9), (2, 1, 3, 4), (1, 15, 2, 2)])
def test_segment_tail_wrapper(K_, B, K, h, w):
    from cutie_b200.utils.tensor_utils import aggregate
    x = 4 * torch.randn(B, K, h, w, generator=torch.Generator().manual_seed(K))
    lg_want = F.interpolate(aggregate(torch.sigmoid(x), dim=1), scale_factor=4, mode='bilinear', align_corners=False)
    lg, pr = K_.segment_tail(x)
    assert torch.allclose(lg, lg_want, rtol=1e-5, atol=1e-5)
    assert torch.allclose(pr, F.softmax(lg_want, dim=1), rtol=1e-5, atol=1e-6)
    with pytest.raises(AssertionError):
        K_.segment_tail(torch.zeros(1, 16, 2, 2))                      # 1 + K channels must fit the kernel's registers


@pytest.mark.parametrize('B,K,C,h,w', [(1, 3, 16, 5, 6), (2, 2, 4, 7, 5), (1, 1, 8, 1, 1)])
def test_upsample2x_add_wrapper(K_, B, K, C, h, w):
    g = torch.Generator().manual_seed(C)
    gg = torch.randn(B, K, C, h, w, generator=g)
    skip = torch.randn(B, C, 2 * h, 2 * w, generator=g)
    want = F.interpolate(gg.flatten(0, 1), scale_factor=2, mode='bilinear', align_corners=False).view(B, K, C, 2 * h, 2 * w) \
        + skip.unsqueeze(1)
    assert torch.allclose(K_.upsample2x_add(gg, skip), want, rtol=1e-5, atol=2e-6)


def test_prob_to_mask_wrapper(K_):
    g = torch.Generator().manual_seed(0)
    prob = torch.rand(4, 9, 13, generator=g)
    lut = torch.tensor([0, 7, 3, 11, 99], dtype=torch.int64)
    want = lut[prob.argmax(0)]
    assert torch.equal(K_.prob_to_mask(prob, lut), want)
    strided = torch.rand(4, 9, 26, generator=g)[:, :, :13]             # row stride != width
    assert torch.equal(K_.prob_to_mask(strided, lut), lut[strided.argmax(0)])


@pytest.mark.parametrize('shape', [(3, 16, 9, 12), (1, 128, 6, 7), (2, 5, 1, 1), (1, 3, 2, 5)])
@pytest.mark.parametrize('relu', [False, True])
def test_conv3x3_c1_wrapper(K_, shape, relu):
    g = torch.Generator().manual_seed(shape[1] + relu)
    x = torch.randn(*shape, generator=g)
    conv = torch.nn.Conv2d(shape[1], 1, 3, padding=1)
    with torch.no_grad():
        want = conv(torch.relu(x) if relu else x)
        got = K_.conv3x3_c1(x, conv.weight, conv.bias, relu_input=relu)
    assert got.shape == want.shape and torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    xs = x.transpose(-1, -2).contiguous().transpose(-1, -2)                 # strided input: the wrapper copies
    with torch.no_grad():
        assert torch.allclose(K_.conv3x3_c1(xs, conv.weight, conv.bias, relu_input=relu), want, rtol=1e-5, atol=1e-5)
