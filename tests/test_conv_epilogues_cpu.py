"""Host logic of fuse.ConvEpilogueFuser: act(conv(x) + bias [+ z]) in the cheapest of three forms ('aten' = PyTorch's
launches, 'cudnn' = cuDNN's fused conv-bias-add-ReLU graph, 'kernel' = bias-less convolution + cutie_bias_act), chosen
per layer by an on-device trial.  No GPU here: the fused op is emulated, cutie_bias_act runs as its CPU emulation
(tests/cpu_kernels.py) and the trial clock is scripted, so what is tested is the wiring -- folded trunks route every
ReLU / residual add / bias through the fuser, the trial accepts, rejects and survives exceptions, and all forms agree
with the un-folded modules."""
import copy

import pytest
import torch
import torch.nn.functional as F

from cutie_b200.model import fuse
from cutie_b200.model.backbone import ResNetTrunk
from cutie_b200.model.blocks import ChannelAttnResBlock, ObjConv2d, ObjResBlock


class _FakeDeviceFuser(fuse.ConvEpilogueFuser):
    """Runs on CPU tensors; `ms` scripts the trial clock per form; `broken` makes the cuDNN form misbehave."""

    def __init__(self, ms=None, broken=None, forms=fuse.ConvEpilogueFuser.FORMS):
        super().__init__(enabled=True, forms=forms)
        self.ms = {**{'aten': 3.0, 'cudnn': 1.0, 'kernel': 2.0, 'stem-aten': 3.0, 'pool': 0.5, 'pool+pad': 0.7}, **(ms or {})}
        self.broken = broken
        self.calls = {'aten': 0, 'cudnn': 0, 'kernel': 0}
        self._which = None

    def _eligible(self, conv, x):
        return self.enabled and conv.bias is not None and x.dim() == 4 and not torch.is_grad_enabled()

    @staticmethod
    def _capturing():
        return False

    def fused(self, conv, x, z=None):
        self.calls['cudnn'] += 1
        self._which = 'cudnn'
        if self.broken == 'raise':
            raise RuntimeError('CUDNN_STATUS_NOT_SUPPORTED')
        y = F.conv2d(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)
        if z is None:
            assert self._zero_like_output(conv, x).shape == y.shape       # the operand cuDNN would be handed
        else:
            y = y + z
        if self.broken == 'wrong':
            y = y + 1.0
        if self.broken == 'nan':
            y = y * float('nan')
        return torch.relu(y)

    def kernel(self, conv, x, z=None, relu=True):
        self.calls['kernel'] += 1
        self._which = 'kernel'
        return super().kernel(conv, x, z, relu)

    def run(self, form, conv, x, z=None, relu=True):
        self._form = form
        return super().run(form, conv, x, z, relu)

    def _stem_run(self, form, conv, x, aten):
        out = super()._stem_run(form, conv, x, aten)
        self._form = form if form != 'aten' else 'stem-aten'
        return out

    def _time(self, fn):
        self._form = 'aten'                 # the reference form is timed through unfused(), not run()
        fn()
        return self.ms.get(self._form, self.ms[self._which])

    def unfused(self, conv, x, z=None, relu=True):
        self.calls['aten'] += 1
        self._which = 'aten'
        return fuse.ConvEpilogueFuser.unfused(conv, x, z, relu)


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
        f = _FakeDeviceFuser()
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
    # every conv with a ReLU behind it took the (scripted-fastest) cuDNN form, the projection shortcuts -- bias only,
    # no ReLU, so no cuDNN form -- took ours
    assert rep['cudnn'] == n - downsamples and rep['kernel'] == downsamples and rep['aten'] == 0 and rep['errors'] == 0
    assert f.calls['cudnn'] - calls_first['cudnn'] == rep['cudnn']       # after the trials: one call per layer
    assert f.calls['kernel'] - calls_first['kernel'] == rep['kernel']
    assert f.calls['aten'] == calls_first['aten']
    assert rep['trial_ms_saved_per_pass'] == pytest.approx(2.0 * rep['cudnn'] + 1.0 * rep['kernel'])


@pytest.mark.parametrize('broken', ['raise', 'wrong', 'nan'])
def test_trial_drops_a_bad_form_and_keeps_the_next_best(broken, cpu_kernels):
    torch.manual_seed(1)
    blk = ChannelAttnResBlock(8, 8).eval()
    x = torch.randn(2, 8, 12, 10)
    with torch.inference_mode():
        ref = blk(x)
        f = _FakeDeviceFuser(broken=broken)
        fuse.attach_epilogue_fuser(blk, f)
        out = blk(x)
        out2 = blk(x)
    assert torch.equal(out, ref) and torch.equal(out2, ref)          # 'kernel' repeats ATen's arithmetic exactly
    rep = f.report()
    # conv1 (+ReLU): cuDNN form rejected -> ours; conv2 (bias only): ours
    assert rep['cudnn'] == 0 and rep['kernel'] == 2 and rep['aten'] == 0 and rep['errors'] == 1 and rep['first_error']


def test_trial_keeps_pytorch_launches_when_they_are_fastest(cpu_kernels):
    torch.manual_seed(2)
    blk = ObjResBlock(6, 4).eval()
    g = torch.randn(1, 3, 6, 9, 7)
    with torch.inference_mode():
        ref = blk(g)
        f = _FakeDeviceFuser(ms=dict(aten=1.0, cudnn=5.0, kernel=4.0))
        fuse.attach_epilogue_fuser(blk, f)
        out = blk(g)
        n_other = f.calls['cudnn'] + f.calls['kernel']
        out = blk(g)
    assert torch.equal(out, ref)
    rep = f.report()
    assert rep['aten'] == 3 and rep['cudnn'] == 0 and rep['kernel'] == 0     # conv1, conv2, 1x1 downsample
    assert f.calls['cudnn'] + f.calls['kernel'] == n_other                     # losers are never called again


def test_kernel_form_handles_object_convs_residuals_and_layouts(cpu_kernels):
    """ObjConv2d (5-D tensors through fold/unfold), residual operand, channels-last storage: 'kernel' == 'aten'."""
    torch.manual_seed(4)
    conv = ObjConv2d(5, 7, 3, padding=1).eval()
    g = torch.randn(2, 3, 5, 6, 8)
    f = _FakeDeviceFuser(forms=('aten', 'kernel'))
    with torch.inference_mode():
        ref = conv(g)
        fuse.attach_epilogue_fuser(conv, f)
        assert torch.equal(conv(g), ref) and f.report()['kernel'] == 1
        plain = torch.nn.Conv2d(5, 7, 3, padding=1).eval()
        x, z = torch.randn(2, 5, 6, 8), torch.randn(2, 7, 6, 8)
        want = torch.relu(plain(x) + z)
        fuse.attach_epilogue_fuser(plain, f)
        assert torch.equal(fuse.conv_add_relu(plain, x, z), want)
        xcl = x.contiguous(memory_format=torch.channels_last)
        got = fuse.conv_add_relu(plain, xcl, z)                     # z in the other storage order
        assert torch.allclose(got, want, atol=1e-6)
    assert f.report()['errors'] == 0


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
    assert torch.equal(out, ref) and not real.decisions and not real.errors


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
    assert net.conv_epilogues.enabled and not net.conv_epilogues.decisions      # CPU tensors: PyTorch's launches
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


# ---- pixel-side glue: ATen chains vs cutie kernels (utils/op_trials.OpTrials) ------------------------------------
class _FakeDeviceTrials:
    """OpTrials on CPU tensors with a scripted clock (kernel_ms, aten_ms per op)."""

    def __new__(cls, ms=None, **kw):
        from cutie_b200.utils.op_trials import OpTrials

        class T(OpTrials):
            def _eligible(self, probe):
                return self.enabled and not torch.is_grad_enabled()

            @staticmethod
            def _capturing():
                return False

            def _time(self, fn):
                fn()
                self._n = getattr(self, '_n', 0) + 1
                return (ms or (1.0, 2.0))[(self._n - 1) % 2]        # kernel first, then ATen (order in _trial)
        return T(**kw)


def test_glue_ops_route_through_trials_and_match(cpu_kernels):
    from cutie_b200.config import default_config
    from cutie_b200.model.cutie import CUTIE
    from cutie_b200.utils.op_trials import attach_op_trials
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
        t = _FakeDeviceTrials()
        attach_op_trials(net, t)
        fused = net.pixel_fusion(pix_feat, pixel, sensory, last_mask)
        sens, logits = net.mask_decoder(ms, fused, sensory)
        summ, _ = net.object_summarizer(last_mask, pixel)
    rep = t.report()
    assert rep['errors'] == 0, rep
    assert rep['ops']['area_pool']['kernel'] >= 3            # mask /16 (fusion + summarizer share a shape), g8 /2, g4 /4
    assert rep['ops']['eca_scale_add']['kernel'] >= 1 and rep['ops']['gated_update']['kernel'] >= 1
    for a, b in ((fused, ref_fused), (sens, ref_sens), (logits, ref_logits), (summ, ref_summ)):
        assert float((a - b).abs().max()) <= 2e-5 * (float(b.abs().max()) + 1e-6)


def test_glue_trial_rejects_a_wrong_kernel_and_propagates_kernel_errors(cpu_kernels, monkeypatch):
    import cutie_b200.kernels as K_
    from cutie_b200.model.blocks import area_resize
    from cutie_b200.utils.op_trials import attach_op_trials
    owner = torch.nn.Identity()
    t = _FakeDeviceTrials()
    attach_op_trials(owner, t)
    x = torch.rand(2, 3, 8, 12)
    with torch.inference_mode():
        want = F.interpolate(x.reshape(-1, 1, 8, 12), size=(2, 3), mode='area').reshape(2, 3, 2, 3)
        monkeypatch.setattr(K_, 'area_pool', lambda x_, f: K_.__dict__['area_pool__orig'](x_, f) + 1.0, raising=False)
        K_.__dict__['area_pool__orig'] = cpu_kernels.area_pool
        got = area_resize(owner, x, (2, 3))
        assert torch.equal(got, want) and t.report()['errors'] == 1 and t.report()['ops']['area_pool'] == {'kernel': 0, 'aten': 1}
        assert torch.equal(area_resize(owner, x, (2, 3)), want)            # decision is sticky

        def boom(x_, f):
            raise K_.KernelError('libcutie_b200.so not found')
        monkeypatch.setattr(K_, 'area_pool', boom)
        with pytest.raises(K_.KernelError):                               # never absorbed into an ATen fallback
            area_resize(owner, torch.rand(1, 8, 8), (2, 2))
        # non-integer ratios stay with PyTorch
        y = torch.rand(1, 9, 10)
        assert torch.equal(area_resize(owner, y, (4, 4)), F.interpolate(y[None], size=(4, 4), mode='area')[0])
    del K_.__dict__['area_pool__orig']


def test_whole_stream_with_every_optional_form_active(cpu_kernels):
    """InferenceCore over a short clip with BN folding, channels-last trunks, the epilogue fuser choosing the cuDNN /
    kernel forms and the glue trials choosing our kernels (all emulated on CPU) against the plain model: the closest
    CPU stand-in for the bench configuration -- every stride / layout hand-over between the forms is exercised."""
    from cutie_b200.config import default_config
    from cutie_b200.inference.inference_core import InferenceCore
    from cutie_b200.model.cutie import CUTIE
    from cutie_b200.utils.op_trials import attach_op_trials
    from oracle.synth import synthetic_state_dict, synthetic_video
    cfg = default_config(mem_every=2, max_mem_frames=3)

    def net():
        n = CUTIE(cfg).eval()
        n.load_state_dict(synthetic_state_dict(n.state_dict(), 0))
        return n
    plain, fast = net(), net().optimize_for_inference()
    f, t = _FakeDeviceFuser(), _FakeDeviceTrials()
    fuse.attach_epilogue_fuser(fast, f)
    attach_op_trials(fast, t)
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
    rf, rt = f.report(), t.report()
    assert rf['errors'] == 0 and rt['errors'] == 0, (rf, rt)
    assert rf['cudnn'] >= 40 and rf['kernel'] >= 20                       # trunks / bias-only convolutions
    assert rf['stem_pool'] == 2                                           # pixel- and mask-encoder stems
    assert set(rt['ops']) == {'area_pool', 'eca_scale_add', 'gated_update', 'qt_p2q_splits', 'caresblock_channels_last',
                              'segment_tail', 'objresblock_channels_last', 'pred_conv3x3'}
    # (on the CPU the bilinear + skip add can hand the second block a non-contiguous tensor, which keeps the NCHW form)
    assert rt['ops']['objresblock_channels_last']['kernel'] >= 1 and rt['ops']['objresblock_channels_last']['aten'] == 0
    assert rt['ops']['caresblock_channels_last'] == {'kernel': 1, 'aten': 0}       # one geometry, three blocks
    assert len(rt['ops']['qt_p2q_splits']['picked']) == 1                  # one decision per (objects, pixels)


def test_pick_times_every_candidate_once_and_sticks():
    t = _FakeDeviceTrials()
    ran = []

    def run(c):
        ran.append(c)
    t._time = lambda fn: (fn(), (5.0, 1.0, 3.0)[len(ran) - 1])[1]      # scripted clock: a -> 5 ms, b -> 1 ms, c -> 3 ms
    probe = torch.zeros(1)
    with torch.inference_mode():
        assert t.pick('op', (1,), ['a', 'b', 'c'], run, probe) == 'b'
        assert ran == ['a', 'b', 'c']
        assert t.pick('op', (1,), ['a', 'b', 'c'], run, probe) == 'b' and ran == ['a', 'b', 'c']
        assert t.pick('op', (2,), ['only'], run, probe) == 'only'
    assert t.report()['ops']['op'] == {'picked': ['b']}


def test_stem_gets_padded_input_forms(cpu_kernels):
    """3- and 5-channel stems: every form is also offered on a zero-padded (4 / 8 channel) input with a zero-padded
    weight twin; the results agree with the plain convolution and the twin shares the bias Parameter."""
    torch.manual_seed(5)
    for cin, cp in ((3, 4), (5, 8)):
        conv = torch.nn.Conv2d(cin, 16, 7, stride=2, padding=3).eval()
        x = torch.randn(2, cin, 20, 28).contiguous(memory_format=torch.channels_last)
        conv = conv.to(memory_format=torch.channels_last)
        f = _FakeDeviceFuser(ms={'aten': 9.0, 'cudnn': 8.0, 'kernel': 7.0, 'aten+pad': 6.0, 'cudnn+pad': 1.0,
                                 'kernel+pad': 5.0})           # scripted clock: the padded cuDNN form is the fastest
        seen = []
        orig_run = f.run

        def run(form, *a, **k):
            seen.append(form)
            return orig_run(form, *a, **k)
        f.run = run
        with torch.inference_mode():
            want = torch.relu(conv(x))
            fuse.attach_epilogue_fuser(conv, f)
            got = fuse.conv_relu(conv, x)
            again = fuse.conv_relu(conv, x)
        assert {'cudnn+pad', 'kernel+pad', 'aten+pad'} <= set(seen)
        assert list(f.decisions.values()) == ['cudnn+pad'] and f.report()['padded_input'] == 1
        tw = f._twins[id(conv)][1]
        assert tw.in_channels == cp and tw.bias is conv.bias and float(tw.weight[:, cin:].abs().sum()) == 0
        assert torch.equal(tw.weight[:, :cin], conv.weight)
        assert tw.weight.is_contiguous(memory_format=torch.channels_last)
        assert torch.allclose(got, want, atol=1e-5) and torch.allclose(again, want, atol=1e-5)
        assert f.report()['errors'] == 0
    # wide convolutions never get padded forms
    assert fuse.ConvEpilogueFuser._padded_channels(torch.nn.Conv2d(64, 64, 3)) == 0
    assert fuse.ConvEpilogueFuser._padded_channels(torch.nn.Conv2d(258, 64, 1)) == 0
