"""Host logic of fuse.ConvEpilogueFuser and utils.dispatch.GlueDispatch: act(conv(x) + bias [+ z]) in the form a
COMMITTED rule names ('cudnn' = cuDNN's fused conv-bias-add-ReLU graph for ReLU epilogues, 'kernel' = bias-less convolution
+ cutie_bias_act otherwise, 'pool' for the ResNet stems), and the pixel-side glue ops per the committed table.  No GPU here:
the fused op is emulated and cutie_bias_act runs as its CPU emulation (tests/cpu_kernels.py); what is tested is the wiring
-- folded trunks route every ReLU / residual add / bias through the fuser, every form agrees with the un-folded modules,
and the choice never depends on anything measured at run time."""
import copy

import pytest
import torch
import torch.nn.functional as F

from cutie_b200.model import fuse
from cutie_b200.model.backbone import ResNetTrunk
from cutie_b200.model.blocks import ChannelAttnResBlock, ObjConv2d, ObjResBlock


EPILOGUE_ONLY = {'relu': 'cudnn', 'linear': 'kernel', 'stem': 'pool'}     # RULE without the tensor-core convolution


class _FakeDeviceFuser(fuse.ConvEpilogueFuser):
    """Runs the device forms on CPU tensors (emulated) and counts the calls per form."""

    def __init__(self, rule=None):
        super().__init__(enabled=True, rule=rule)
        self.calls = {'aten': 0, 'cudnn': 0, 'kernel': 0}

    def _eligible(self, conv, x):
        return self.enabled and conv.bias is not None and x.dim() == 4 and not torch.is_grad_enabled()

    def fused(self, conv, x, z=None):
        self.calls['cudnn'] += 1
        y = F.conv2d(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)
        if z is None:
            assert self._zero_like_output(conv, x).shape == y.shape       # the operand cuDNN would be handed
        else:
            y = y + z
        return torch.relu(y)

    def kernel(self, conv, x, z=None, relu=True):
        self.calls['kernel'] += 1
        return super().kernel(conv, x, z, relu)

    def unfused(self, conv, x, z=None, relu=True, relu_in=False):
        self.calls['aten'] += 1
        return fuse.ConvEpilogueFuser.unfused(conv, x, z, relu, relu_in)


def _randomise_bn(m):
    g = torch.Generator().manual_seed(0)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data = 0.5 + torch.rand(mod.weight.shape, generator=g)
            mod.bias.data = torch.randn(mod.bias.shape, generator=g) * 0.1
            mod.running_mean = torch.randn(mod.running_mean.shape, generator=g) * 0.1
            mod.running_var = 0.5 + torch.rand(mod.running_var.shape, generator=g)


class _Trunk(torch.nn.Module):
    def __init__(self, arch):
        super().__init__()
        t = ResNetTrunk(arch)
        self.conv1, self.bn1, self.layer1, self.layer2, self.layer3 = t.conv1, t.bn1, t.layer1, t.layer2, t.layer3

    def forward(self, x):
        if getattr(self, 'bn_folded', False):
            x = fuse.conv_relu(self.conv1, x)
        else:
            x = F.relu(self.bn1(self.conv1(x)))
        return self.layer3(self.layer2(self.layer1(F.max_pool2d(x, 3, stride=2, padding=1))))


@pytest.mark.parametrize('arch', ['resnet18', 'resnet50'])
def test_folded_trunk_routes_through_fuser_and_matches(arch, cpu_kernels):
    torch.manual_seed(0)
    net = _Trunk(arch).eval()
    _randomise_bn(net)
    x = torch.randn(2, 3, 48, 64)
    with torch.inference_mode():
        ref = net(x)
        n = fuse.fold_trunk_(net)
        assert net.bn_folded and all(getattr(u, 'bn_folded', False) for s in (net.layer1, net.layer2, net.layer3) for u in s)
        off = net(x)                      # no fuser attached: convolution + add + clamp, as before
        f = _FakeDeviceFuser(rule=EPILOGUE_ONLY)          # (trunks are channels-last on the device: never 'tc')
        assert fuse.attach_epilogue_fuser(net, f) == n
        on = net(x)
        calls_first = dict(f.calls)
        on2 = net(x)
    scale = float(ref.abs().max())
    assert float((off - ref).abs().max()) < 2e-5 * scale
    assert float((on - ref).abs().max()) < 2e-5 * scale
    assert torch.equal(on, on2)
    rep = f.report()
    downsamples = 3 if arch == 'resnet50' else 2
    # every conv with a ReLU behind it takes the cuDNN form, the projection shortcuts (bias only, no ReLU) take ours
    assert rep['layers'] == {'cudnn': n - downsamples, 'kernel': downsamples}
    assert f.calls['aten'] == 0
    assert f.calls['cudnn'] == 2 * calls_first['cudnn'] and f.calls['kernel'] == 2 * calls_first['kernel']   # same route twice


def test_rule_is_the_only_input_of_the_choice(cpu_kernels):
    """Another rule routes the same block differently; results agree ('kernel' repeats ATen's arithmetic exactly)."""
    torch.manual_seed(1)
    blk = ChannelAttnResBlock(8, 8).eval()
    x = torch.randn(2, 8, 12, 10)
    with torch.inference_mode():
        ref = blk(x)
        f = _FakeDeviceFuser(rule={'relu': 'kernel', 'linear': 'kernel', 'stem': 'pool'})
        fuse.attach_epilogue_fuser(blk, f)
        out = blk(x)
        g = _FakeDeviceFuser(rule={'relu': 'aten', 'linear': 'aten', 'stem': 'aten'})
        fuse.attach_epilogue_fuser(blk, g)
        out2 = blk(x)
    assert torch.equal(out, ref) and torch.equal(out2, ref)
    assert f.report()['layers'] == {'kernel': 2} and f.calls['cudnn'] == 0
    assert g.report()['layers'] == {'aten': 2} and g.calls['kernel'] == 0


def test_eligible_3x3_layers_take_the_tensor_core_form(cpu_kernels):
    """3x3 / stride 1 / pad 1 with Cin % 32 == 0 and Cout % 128 == 0 on dense NCHW: the convolution itself goes to
    cutie_conv_tc with the input ReLU, bias, residual and output ReLU inside (emulated here); the operand image is built
    once per weight version; other geometries keep the cuDNN forms."""
    torch.manual_seed(5)
    blk = ObjResBlock(32, 128).eval()                       # conv1 32 -> 128 (tc), conv2 128 -> 128 (tc), 1x1 shortcut (kernel)
    g = torch.randn(1, 2, 32, 6, 5)
    with torch.inference_mode():
        ref = blk(g)
        f = _FakeDeviceFuser()
        fuse.attach_epilogue_fuser(blk, f)
        out = blk(g)
        out2 = blk(g)
    assert torch.allclose(out, ref, atol=1e-5) and torch.equal(out, out2)
    assert f.report()['layers'] == {'tc': 3}                # (the 1x1 projection shortcut too)
    assert len(f._images) == 3
    img_before = f._images[id(blk.conv1)][1]
    with torch.no_grad():
        blk.conv1.weight.mul_(2.0)                          # an in-place write bumps the version: the image is rebuilt
    with torch.inference_mode():
        out3 = blk(g)
    assert f._images[id(blk.conv1)][1] is not img_before and not torch.allclose(out3, out)
    car = ChannelAttnResBlock(128, 128).eval()
    x = torch.randn(2, 128, 5, 4)
    with torch.inference_mode():
        want = car(x)
        h = _FakeDeviceFuser()
        fuse.attach_epilogue_fuser(car, h)
        assert torch.allclose(car(x), want, atol=1e-5) and h.report()['layers'] == {'tc': 2}
        k = _FakeDeviceFuser()
        fuse.attach_epilogue_fuser(car, k)
        xcl = x.contiguous(memory_format=torch.channels_last)
        assert torch.allclose(car(xcl), want, atol=1e-5) and k.report()['layers'] == {'tc': 2}     # either memory format
        off = _FakeDeviceFuser(rule={'relu': 'cudnn', 'linear': 'kernel', 'stem': 'pool'})     # a rule without 'conv'
        fuse.attach_epilogue_fuser(car, off)
        assert torch.allclose(car(x), want, atol=1e-5) and 'tc' not in off.report()['layers']


def test_object_resblock_residual_rides_in_the_epilogue(cpu_kernels):
    torch.manual_seed(2)
    blk = ObjResBlock(6, 4).eval()
    g = torch.randn(1, 3, 6, 9, 7)
    with torch.inference_mode():
        ref = blk(g)
        f = _FakeDeviceFuser()
        fuse.attach_epilogue_fuser(blk, f)
        out = blk(g)
    assert torch.allclose(out, ref, atol=1e-6)
    assert f.report()['layers'] == {'cudnn': 1, 'kernel': 2}                # conv1 (+ReLU); conv2 + residual, 1x1 shortcut


def test_kernel_form_handles_object_convs_residuals_and_layouts(cpu_kernels):
    """ObjConv2d (5-D tensors through fold/unfold), residual operand, channels-last storage: 'kernel' == 'aten'."""
    torch.manual_seed(4)
    conv = ObjConv2d(5, 7, 3, padding=1).eval()
    g = torch.randn(2, 3, 5, 6, 8)
    f = _FakeDeviceFuser(rule={'relu': 'kernel', 'linear': 'kernel', 'stem': 'pool'})
    with torch.inference_mode():
        ref = conv(g)
        fuse.attach_epilogue_fuser(conv, f)
        assert torch.equal(conv(g), ref) and f.report()['layers'] == {'kernel': 1}
        plain = torch.nn.Conv2d(5, 7, 3, padding=1).eval()
        x, z = torch.randn(2, 5, 6, 8), torch.randn(2, 7, 6, 8)
        want = torch.relu(plain(x) + z)
        fuse.attach_epilogue_fuser(plain, f)
        assert torch.equal(fuse.conv_add_relu(plain, x, z), want)
        xcl = x.contiguous(memory_format=torch.channels_last)
        got = fuse.conv_add_relu(plain, xcl, z)                     # z in the other storage order
        assert torch.allclose(got, want, atol=1e-6)


def test_cpu_tensors_never_reach_the_device_forms():
    """The real fuser (enabled) must leave CPU tensors on PyTorch's launches: the oracle harness borrows these modules
    on CPU, and neither torch.cudnn_convolution_add_relu nor cutie_bias_act exists there."""
    torch.manual_seed(3)
    blk = ChannelAttnResBlock(8, 8).eval()
    x = torch.randn(1, 8, 9, 9)
    with torch.inference_mode():
        ref = blk(x)
        real = fuse.ConvEpilogueFuser(enabled=True)
        fuse.attach_epilogue_fuser(blk, real)
        out = blk(x)
    assert torch.equal(out, ref) and not real.counts


def test_zero_operand_geometry():
    f = fuse.ConvEpilogueFuser()
    conv = torch.nn.Conv2d(3, 5, 7, stride=2, padding=3)
    x = torch.randn(2, 3, 33, 47)
    z = f._zero_like_output(conv, x)
    assert z.shape == conv(x).shape and float(z.abs().sum()) == 0
    assert f._zero_like_output(conv, x) is z                                # cached per geometry
    xcl = x.contiguous(memory_format=torch.channels_last)
    zcl = f._zero_like_output(conv, xcl)
    assert zcl.is_contiguous(memory_format=torch.channels_last) and zcl is not z


def test_optimize_for_inference_attaches_one_fuser_per_model():
    from cutie_b200.config import default_config
    from cutie_b200.model.cutie import CUTIE
    cfg = default_config()
    torch.manual_seed(0)
    net = CUTIE(cfg).eval()
    _randomise_bn(net)
    x = torch.randn(1, 3, 64, 96)
    with torch.inference_mode():
        ref = net.pixel_encoder(x)
        net.optimize_for_inference()
        out = net.pixel_encoder(x)
    convs = [m for m in net.modules() if isinstance(m, torch.nn.Conv2d)]
    assert convs and all(m.epilogue_fuser is net.conv_epilogues for m in convs)
    assert net.conv_epilogues.enabled and not net.conv_epilogues.counts         # CPU tensors: PyTorch's launches
    assert not any('epilogue' in k or '_conv_forward' in k for k in net.state_dict())
    for a, b in zip(out, ref):
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max())
    other = CUTIE(cfg).eval()                                                    # an un-optimised model is untouched
    assert not hasattr(other, 'conv_epilogues')
    assert all('epilogue_fuser' not in m.__dict__ and '_conv_forward' not in m.__dict__ for m in other.modules())
    # a deep copy points at its own convolutions and its own (empty) fuser
    dup = copy.deepcopy(net.key_proj)
    for m in dup.modules():
        if isinstance(m, torch.nn.Conv2d):
            assert m._conv_forward.conv is m and m._conv_forward.fuser is m.epilogue_fuser
            assert m.epilogue_fuser is not net.conv_epilogues
    with torch.inference_mode():
        f16 = torch.randn(1, 1024, 4, 6)
        a = net.key_proj(f16, need_s=True, need_e=True)
        b = dup(f16, need_s=True, need_e=True)
    assert all(torch.equal(p, q) for p, q in zip(a, b))


# ---- pixel-side glue: ATen chains vs cutie kernels (utils/dispatch.GlueDispatch) ------------------------------------
def _fake_device_dispatch(**kw):
    """GlueDispatch that treats CPU tensors as eligible (the kernels are emulated on CPU by the cpu_kernels fixture)."""
    from cutie_b200.utils.dispatch import GlueDispatch

    class T(GlueDispatch):
        def _eligible(self, probe):
            return self.enabled and not torch.is_grad_enabled()
    return T(**kw)


def test_glue_ops_route_through_the_table_and_match(cpu_kernels):
    from cutie_b200.config import default_config
    from cutie_b200.model.cutie import CUTIE
    from cutie_b200.utils.dispatch import attach_glue_dispatch
    cfg = default_config()
    torch.manual_seed(0)
    net = CUTIE(cfg).eval()
    g = torch.Generator().manual_seed(1)
    B, K, h, w = 1, 2, 4, 6
    pix_feat = torch.randn(B, 256, h, w, generator=g)
    pixel = torch.randn(B, K, 256, h, w, generator=g)
    sensory = torch.randn(B, K, 256, h, w, generator=g)
    last_mask = torch.rand(B, K, 16 * h, 16 * w, generator=g)
    ms = [torch.randn(B, 1024, h, w, generator=g), torch.randn(B, 512, 2 * h, 2 * w, generator=g),
          torch.randn(B, 256, 4 * h, 4 * w, generator=g)]
    with torch.inference_mode():
        ref_fused = net.pixel_fusion(pix_feat, pixel, sensory, last_mask)
        # the decoder's UpsampleBlock needs the CUDA kernel for CUDA tensors only; on CPU it is plain PyTorch
        ref_sens, ref_logits = net.mask_decoder(ms, ref_fused, sensory)
        ref_summ, _ = net.object_summarizer(last_mask, pixel)
        t = _fake_device_dispatch()
        attach_glue_dispatch(net, t)
        fused = net.pixel_fusion(pix_feat, pixel, sensory, last_mask)
        sens, logits = net.mask_decoder(ms, fused, sensory)
        summ, _ = net.object_summarizer(last_mask, pixel)
    assert t.calls['area_pool'] >= 3            # mask /16 (fusion + summarizer), g8 /2, g4 /4
    assert t.calls['eca_scale_add'] >= 1 and t.calls['gated_update'] >= 1
    for a, b in ((fused, ref_fused), (sens, ref_sens), (logits, ref_logits), (summ, ref_summ)):
        assert float((a - b).abs().max()) <= 2e-5 * (float(b.abs().max()) + 1e-6)


def test_glue_dispatch_is_a_table_lookup_and_never_absorbs_kernel_errors(cpu_kernels, monkeypatch):
    import cutie_b200.kernels as K_
    from cutie_b200.model.blocks import area_resize
    from cutie_b200.utils.dispatch import GLUE_TABLE, attach_glue_dispatch
    assert set(GLUE_TABLE) == {'area_pool', 'eca_scale_add', 'gated_update', 'segment_tail', 'pred_conv3x3',
                               'caresblock_channels_last', 'objresblock_channels_last'}
    owner = torch.nn.Identity()
    x = torch.rand(2, 3, 8, 12)
    with torch.inference_mode():
        want = F.interpolate(x.reshape(-1, 1, 8, 12), size=(2, 3), mode='area').reshape(2, 3, 2, 3)
        off = _fake_device_dispatch(table={**GLUE_TABLE, 'area_pool': False})
        attach_glue_dispatch(owner, off)
        assert torch.equal(area_resize(owner, x, (2, 3)), want) and not off.calls      # table says ATen: kernel untouched
        on = _fake_device_dispatch()
        attach_glue_dispatch(owner, on)
        assert torch.allclose(area_resize(owner, x, (2, 3)), want, atol=1e-6) and on.calls == {'area_pool': 1}

        def boom(x_, f):
            raise K_.KernelError('libcutie_b200.so not found')
        monkeypatch.setattr(K_, 'area_pool', boom)
        with pytest.raises(K_.KernelError):                               # never absorbed into an ATen fallback
            area_resize(owner, torch.rand(1, 8, 8), (2, 2))
        # non-integer ratios stay with PyTorch
        y = torch.rand(1, 9, 10)
        assert torch.equal(area_resize(owner, y, (4, 4)), F.interpolate(y[None], size=(4, 4), mode='area')[0])


def test_whole_stream_with_every_optional_form_active(cpu_kernels):
    """InferenceCore over a short clip with BN folding, channels-last trunks, the epilogue fuser routing to the cuDNN /
    kernel forms and every glue op on our kernels -- including the two channels-last block variants the committed table
    leaves off (all emulated on CPU) -- against the plain model: the closest CPU stand-in for the bench configuration;
    every stride / layout hand-over between the forms is exercised."""
    from cutie_b200.config import default_config
    from cutie_b200.inference.inference_core import InferenceCore
    from cutie_b200.model.cutie import CUTIE
    from cutie_b200.utils.dispatch import GLUE_TABLE, attach_glue_dispatch
    from oracle.synth import synthetic_state_dict, synthetic_video
    cfg = default_config(mem_every=2, max_mem_frames=3)

    def net():
        n = CUTIE(cfg).eval()
        n.load_state_dict(synthetic_state_dict(n.state_dict(), 0))
        return n
    plain, fast = net(), net().optimize_for_inference()
    f, t = _FakeDeviceFuser(), _fake_device_dispatch(table={k: True for k in GLUE_TABLE})
    fuse.attach_epilogue_fuser(fast, f)
    attach_glue_dispatch(fast, t)
    a, b = InferenceCore(plain, cfg=cfg), InferenceCore(fast, cfg=cfg)
    frames, mask = synthetic_video(4, 96, 160, 3, seed=3)
    with torch.inference_mode():
        for ti in range(4):
            if ti == 0:
                a.step(frames[0], mask, objects=[1, 2, 3]); b.step(frames[0], mask, objects=[1, 2, 3])
            else:
                pa, pb = a.step(frames[ti]), b.step(frames[ti])
                assert float((a.last_logits - b.last_logits).abs().max()) < 1e-3
                assert float((pa - pb).abs().max()) < 1e-3
    rf = f.report()['layers']
    print(rf)
    # 3x3 / 1x1 layers on the tensor-core form; what stays with the library: the stems ('pool') and the layers whose
    # channel counts do not fit (Cin % 32, Cout < 64: 'kernel' = bias-less cuDNN convolution + cutie_bias_act)
    assert rf['tc'] >= 78 and rf.get('cudnn', 0) <= 4 and rf['kernel'] >= 2
    assert rf['pool'] == 2 and 'aten' not in rf                           # pixel- and mask-encoder stems
    assert set(t.calls) == set(GLUE_TABLE)
