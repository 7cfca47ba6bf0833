"""Host logic of fuse.ConvEpilogueFuser (conv + bias [+ residual] + ReLU as one cuDNN call, chosen per layer by
an on-device trial).  No GPU here: the fused op is emulated and the trial clock is scripted, so what is tested
is the wiring -- folded trunks route every ReLU / residual add through the fuser, the trial accepts, rejects and
survives exceptions, and all forms agree with the un-folded modules."""
import pytest
import torch
import torch.nn.functional as F

from cutie_b200.model import fuse
from cutie_b200.model.backbone import ResNetTrunk
from cutie_b200.model.blocks import ChannelAttnResBlock, ObjResBlock


class _FakeDeviceFuser(fuse.ConvEpilogueFuser):
    def __init__(self, fused_ms=1.0, unfused_ms=3.0, broken=None):
        super().__init__(enabled=True)
        self.ms = {'fused': fused_ms, 'unfused': unfused_ms}
        self.broken = broken
        self.calls = {'fused': 0, 'unfused': 0}
        self._which = None

    def _eligible(self, conv, x):
        return self.enabled and conv.bias is not None and x.dim() == 4 and not torch.is_grad_enabled()

    @staticmethod
    def _capturing():
        return False

    def fused(self, conv, x, z=None):
        self.calls['fused'] += 1
        self._which = 'fused'
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

    def _time(self, fn):
        fn()
        return self.ms[self._which]

    def unfused(self, conv, x, z=None):
        self.calls['unfused'] += 1
        self._which = 'unfused'
        return fuse.ConvEpilogueFuser.unfused(conv, x, z)


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
def test_folded_trunk_routes_through_fuser_and_matches(arch):
    torch.manual_seed(0)
    net = _Trunk(arch).eval()
    _randomise_bn(net)
    x = torch.randn(2, 3, 48, 64)
    with torch.inference_mode():
        ref = net(x)
        n = fuse.fold_trunk_(net)
        assert net.bn_folded and all(getattr(u, 'bn_folded', False) for s in (net.layer1, net.layer2, net.layer3) for u in s)
        # no fuser attached: the folded forward is convolution + add + clamp, as before
        off = net(x)
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
    downsamples = 3 if arch == 'resnet50' else 2         # projection shortcuts keep a plain convolution
    assert rep['fused'] == n - downsamples and rep['three_launch'] == 0 and rep['errors'] == 0
    # after the trials only the fused form runs: one call per decided layer, none of the three-launch form
    assert f.calls['fused'] - calls_first['fused'] == rep['fused']
    assert f.calls['unfused'] == calls_first['unfused']
    assert rep['trial_ms_saved_per_pass'] == pytest.approx(2.0 * rep['fused'])


@pytest.mark.parametrize('broken', ['raise', 'wrong', 'nan'])
def test_trial_rejects_a_bad_fused_engine(broken):
    torch.manual_seed(1)
    blk = ChannelAttnResBlock(8, 8).eval()
    x = torch.randn(2, 8, 12, 10)
    with torch.inference_mode():
        ref = blk(x)
        f = _FakeDeviceFuser(broken=broken)
        fuse.attach_epilogue_fuser(blk, f)
        out = blk(x)
        out2 = blk(x)
    assert torch.equal(out, ref) and torch.equal(out2, ref)
    rep = f.report()
    assert rep['fused'] == 0 and rep['three_launch'] == 1 and rep['errors'] == 1 and rep['first_error']


def test_trial_keeps_three_launches_when_they_are_faster():
    torch.manual_seed(2)
    blk = ObjResBlock(6, 4).eval()
    g = torch.randn(1, 3, 6, 9, 7)
    with torch.inference_mode():
        ref = blk(g)
        f = _FakeDeviceFuser(fused_ms=5.0, unfused_ms=3.0)
        fuse.attach_epilogue_fuser(blk, f)
        out = blk(g)
        n_fused = f.calls['fused']
        out = blk(g)
    assert torch.allclose(out, ref, atol=1e-6)
    assert f.report()['fused'] == 0 and f.calls['fused'] == n_fused      # never called again after losing the trial


def test_cpu_tensors_never_reach_the_fused_op():
    """The real fuser (enabled) must leave CPU tensors on the three-launch form: the oracle harness borrows these
    modules on CPU, and torch.cudnn_convolution_add_relu does not exist there."""
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
    keys_before = set(net.state_dict().keys())
    x = torch.randn(1, 3, 64, 96)
    with torch.inference_mode():
        ref = net.pixel_encoder(x)
        net.optimize_for_inference()
        out = net.pixel_encoder(x)
    convs = [m for m in net.modules() if isinstance(m, torch.nn.Conv2d)]
    assert convs and all(m.epilogue_fuser is net.conv_epilogues for m in convs)
    assert net.conv_epilogues.enabled and not net.conv_epilogues.decisions      # CPU tensors: three launches
    assert 'conv_epilogues' not in net.state_dict() and not any('epilogue' in k for k in net.state_dict())
    for a, b in zip(out, ref):
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max())
    other = CUTIE(cfg).eval()                                                    # an un-optimised model is untouched
    assert not hasattr(other, 'conv_epilogues')
    assert all(not hasattr(m, 'epilogue_fuser') for m in other.modules())
    assert keys_before  # (state_dict of the trunks changes by design: bn.* keys fold into conv bias)
