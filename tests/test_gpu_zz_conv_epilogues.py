"""Convolution epilogues (cutie_b200/model/fuse.ConvEpilogueFuser): bias (+ residual) (+ ReLU) through cuDNN's fused
graph or through cutie_bias_act after a bias-less convolution, against PyTorch's own launches, on the GPU.  The
convolutions are PyTorch/cuDNN stages either side of the hot path (kept as library calls); what is asserted is that
switching the epilogue form changes nothing beyond fp32 rounding, in the form the committed rule names for each layer, and that
cutie_bias_act itself is bit-identical to the ATen ops it replaces.  (The file name sorts last on purpose.)"""
import pytest
import torch

# first run of these cases is at round end: never let one of them wedge the suite (thread method: the process exits)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method='thread')]


def _net(cfg, **opt):
    from cutie_b200.model.cutie import CUTIE
    from oracle.synth import synthetic_state_dict
    net = CUTIE(cfg).eval()
    net.load_state_dict(synthetic_state_dict(net.state_dict(), 0))
    return net.cuda().optimize_for_inference(**opt)


def test_trunks_fused_epilogues_match_three_launches():
    from cutie_b200.config import default_config
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = default_config()
    net = _net(cfg, fuse_epilogues=True)
    g = torch.Generator().manual_seed(5)
    img = torch.randn(1, 3, 240, 432, generator=g).cuda()
    with torch.inference_mode():
        fz = net.conv_epilogues
        out_f = net.pixel_encoder(img)
        out_f2 = net.pixel_encoder(img)
        fz.enabled = False
        out_u = net.pixel_encoder(img)
        fz.enabled = True
    rep = fz.report()
    print('conv epilogues (pixel encoder, 240p, fp32):', rep)
    assert sum(rep['layers'].values()) >= 43 and 'aten' not in rep['layers']   # ResNet-50 stages 1-3: stem + 39 convs + 3 shortcuts
    for a, b, c in zip(out_f, out_f2, out_u):
        scale = float(c.abs().max())
        assert torch.isfinite(a).all()
        assert float((a - c).abs().max()) <= 1e-4 * scale, (float((a - c).abs().max()), scale)
        assert float((b - c).abs().max()) <= 1e-4 * scale


def test_stream_with_fused_epilogues_matches_three_launches():
    """Two frames (one memory frame, one propagated frame through the CUDA graphs): logits with fused epilogues
    vs the same optimised model with the fuser switched off."""
    from cutie_b200.config import default_config
    from cutie_b200.inference.inference_core import InferenceCore
    from oracle.synth import synthetic_video
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = default_config(mem_every=2, max_mem_frames=3)
    on, off = _net(cfg, fuse_epilogues=True), _net(cfg, fuse_epilogues=False)
    a, b = InferenceCore(on, cfg=cfg, use_cuda_graphs=True), InferenceCore(off, cfg=cfg, use_cuda_graphs=True)
    frames, mask = synthetic_video(3, 96, 160, 3, seed=3)
    with torch.inference_mode():
        for ti in range(3):
            x = frames[ti].cuda()
            if ti == 0:
                a.step(x, mask.cuda(), objects=[1, 2, 3]); b.step(x, mask.cuda(), objects=[1, 2, 3])
            else:
                pa, pb = a.step(x), b.step(x)
                assert float((a.last_logits - b.last_logits).abs().max()) < 1e-3
                assert float((pa - pb).abs().max()) < 1e-3
    print('conv epilogues (stream, 96x160):', on.conv_epilogues.report())
    assert not off.conv_epilogues.counts


@pytest.mark.parametrize('shape', [(3, 256, 30, 54), (1, 1, 30, 54), (2, 7, 5, 3), (3, 128, 120, 216), (1, 64, 9, 11)])
@pytest.mark.parametrize('cl', [False, True])
@pytest.mark.parametrize('relu', [False, True])
@pytest.mark.parametrize('with_z', [False, True])
def test_bias_act_kernel_is_bit_identical_to_aten(shape, cl, relu, with_z):
    """cutie_bias_act (vector and scalar paths, NCHW and channels-last, odd sizes) vs y.add_(bias).add_(z).relu_()."""
    import cutie_b200.kernels as K_
    g = torch.Generator().manual_seed(sum(shape) + 2 * cl + relu)
    fmt = torch.channels_last if cl else torch.contiguous_format
    y = torch.randn(*shape, generator=g).cuda().contiguous(memory_format=fmt)
    z = torch.randn(*shape, generator=g).cuda() if with_z else None          # NCHW: exercises the layout copy when cl
    bias = torch.randn(shape[1], generator=g).cuda()
    want = y.clone() + bias.view(1, -1, 1, 1)
    if z is not None:
        want = want + z
    if relu:
        want = torch.relu(want)
    got = K_.bias_act_(y.clone(memory_format=torch.preserve_format), bias, z, relu)
    torch.cuda.synchronize()
    assert got.stride() == y.stride()
    assert torch.equal(got, want)


def test_bias_act_misaligned_views_take_the_scalar_path():
    import cutie_b200.kernels as K_
    base = torch.randn(1 + 2 * 8 * 6 * 4).cuda()
    y = base[1:].view(2, 8, 6, 4)                       # 4-byte aligned only
    bias = torch.randn(8).cuda()
    want = torch.relu(y + bias.view(1, -1, 1, 1))
    got = K_.bias_act_(y.clone()[:], bias, None, True)  # clone is aligned: vector path
    assert torch.equal(got, want)
    got2 = K_.bias_act_(y, bias, None, True)            # in place on the misaligned view
    torch.cuda.synchronize()
    assert torch.equal(got2, want) and got2.data_ptr() == y.data_ptr()


@pytest.mark.parametrize('shape,f', [((1, 3, 480, 864), 16), ((3, 256, 60, 108), 2), ((3, 257, 120, 216), 4), ((2, 5, 9, 12), 3),
                                     ((1, 2, 96, 160), 16)])
def test_area_pool_kernel_matches_interpolate_area(shape, f):
    import torch.nn.functional as F
    import cutie_b200.kernels as K_
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(f)).cuda()
    got = K_.area_pool(x, f)
    want = F.interpolate(x, scale_factor=1.0 / f, mode='area')
    assert got.shape == want.shape and float((got - want).abs().max()) <= 1e-6


@pytest.mark.parametrize('shape', [(3, 256, 30, 54), (2, 8, 5, 3), (1, 64, 9, 11)])
@pytest.mark.parametrize('cl', [False, True])
def test_eca_scale_add_kernel_matches_aten(shape, cl):
    import cutie_b200.kernels as K_
    g = torch.Generator().manual_seed(shape[1])
    fmt = torch.channels_last if cl else torch.contiguous_format
    y = torch.randn(*shape, generator=g).cuda().contiguous(memory_format=fmt)
    x = torch.randn(*shape, generator=g).cuda()                     # NCHW: exercises the layout copy when cl
    conv = torch.nn.Conv1d(1, 1, 5, padding=2, bias=False).cuda()
    with torch.inference_mode():
        gate = conv(y.mean(dim=(2, 3)).unsqueeze(1)).sigmoid().transpose(1, 2).unsqueeze(-1)
        want = y * gate + x
        got = K_.eca_scale_add_(y.clone(memory_format=torch.preserve_format), x, conv.weight)
    assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())


@pytest.mark.parametrize('shape', [(1, 3, 256, 30, 54), (2, 2, 4, 5, 3)])
def test_gated_update_kernel_matches_aten(shape):
    import cutie_b200.kernels as K_
    from cutie_b200.model.blocks import gated_update
    B, K, d, H, W = shape
    g = torch.Generator().manual_seed(d)
    h = torch.randn(*shape, generator=g).cuda()
    v = (2 * torch.randn(B, K, 3 * d, H, W, generator=g)).cuda()
    want = gated_update(h, v)                                        # no owner: the ATen composition
    got = K_.gated_update(h, v)
    assert float((got - want).abs().max()) <= 2e-6


def test_stream_with_glue_kernels_matches_aten_chains():
    """Same as the epilogue stream test, for the glue ops: optimised model with fuse_glue on vs off."""
    from cutie_b200.config import default_config
    from cutie_b200.inference.inference_core import InferenceCore
    from oracle.synth import synthetic_video
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = default_config(mem_every=2, max_mem_frames=3)
    on, off = _net(cfg, fuse_epilogues=False, fuse_glue=True), _net(cfg, fuse_epilogues=False, fuse_glue=False)
    a, b = InferenceCore(on, cfg=cfg, use_cuda_graphs=True), InferenceCore(off, cfg=cfg, use_cuda_graphs=True)
    frames, mask = synthetic_video(3, 96, 160, 3, seed=3)
    with torch.inference_mode():
        for ti in range(3):
            x = frames[ti].cuda()
            if ti == 0:
                a.step(x, mask.cuda(), objects=[1, 2, 3]); b.step(x, mask.cuda(), objects=[1, 2, 3])
            else:
                pa, pb = a.step(x), b.step(x)
                assert float((a.last_logits - b.last_logits).abs().max()) < 1e-3
                assert float((pa - pb).abs().max()) < 1e-3
    rep = on.glue_dispatch.report()
    print('glue ops (stream, 96x160):', rep)
    assert on.glue_dispatch.calls and not off.glue_dispatch.calls


@pytest.mark.parametrize('shape', [(1, 64, 240, 432), (3, 64, 48, 80), (2, 8, 9, 12), (1, 6, 7, 7), (1, 4, 1, 1)])
@pytest.mark.parametrize('cl', [False, True])
def test_bias_relu_maxpool_kernel_is_bit_identical(shape, cl):
    """cutie_bias_relu_maxpool(y, b) == max_pool2d(relu(y + b), 3, 2, 1), NCHW and channels-last (C % 4 != 0 falls back
    to an NCHW copy inside the wrapper)."""
    import torch.nn.functional as F
    import cutie_b200.kernels as K_
    g = torch.Generator().manual_seed(shape[2])
    y = torch.randn(*shape, generator=g).cuda()
    if cl:
        y = y.contiguous(memory_format=torch.channels_last)
    bias = torch.randn(shape[1], generator=g).cuda()
    want = F.max_pool2d(torch.relu(y + bias.view(1, -1, 1, 1)), 3, stride=2, padding=1)
    got = K_.bias_relu_maxpool(y, bias)
    assert got.shape == want.shape and torch.equal(got, want)


def test_pixel_ffn_channels_last_variant_matches_nchw():
    """ChannelAttnResBlock with channels-last weight twins (one layout copy in, channels-last in between) vs the NCHW
    form, both on pure PyTorch/cuDNN launches and with the epilogue fuser and the channels-last entry of the glue table on."""
    from cutie_b200.model.blocks import ChannelAttnResBlock
    from cutie_b200.model.fuse import ConvEpilogueFuser, attach_epilogue_fuser
    from cutie_b200.utils.dispatch import GLUE_TABLE, GlueDispatch, attach_glue_dispatch
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    blk = ChannelAttnResBlock(256, 256).cuda().eval()
    x = torch.randn(3, 256, 30, 54, device='cuda')
    with torch.inference_mode():
        ref = blk(x)
        twins = blk.make_channels_last_twins()
        got = blk._forward(x.contiguous(memory_format=torch.channels_last), *twins)
        assert float((got - ref).abs().max()) <= 2e-4 * float(ref.abs().max())
        fz, tr = ConvEpilogueFuser(), GlueDispatch(table={**GLUE_TABLE, 'caresblock_channels_last': True})
        attach_epilogue_fuser(blk, fz)
        for tw in twins:
            attach_epilogue_fuser(tw, fz)
        attach_glue_dispatch(blk, tr)
        out1, out2 = blk(x), blk(x)
    assert float((out1 - ref).abs().max()) <= 2e-4 * float(ref.abs().max())
    assert float((out2 - ref).abs().max()) <= 2e-4 * float(ref.abs().max())
    print('PixelFFN block dispatch:', tr.calls, fz.report())
    assert tr.calls.get('caresblock_channels_last') == 2


@pytest.mark.parametrize('B,K,h,w', [(1, 3, 120, 216), (2, 1, 24, 40), (1, 15, 6, 10)])
def test_segment_tail_kernel_matches_aten(B, K, h, w):
    import torch.nn.functional as F
    import cutie_b200.kernels as K_
    from cutie_b200.utils.tensor_utils import aggregate
    x = (4 * torch.randn(B, K, h, w, generator=torch.Generator().manual_seed(K))).cuda()
    x[0, 0, 0, 0], x[0, K - 1, -1, -1] = 40.0, -40.0
    lg_want = F.interpolate(aggregate(torch.sigmoid(x), dim=1), scale_factor=4, mode='bilinear', align_corners=False)
    pr_want = F.softmax(lg_want, dim=1)
    lg, pr = K_.segment_tail(x)
    assert float((lg - lg_want).abs().max()) <= 1e-4 and float((pr - pr_want).abs().max()) <= 1e-5


@pytest.mark.parametrize('optimised', [False, True])
def test_encoder_lookahead_gives_the_same_stream(optimised):
    """step(image, next_image=...) runs the next frame's encoder graph on a side stream (two alternating capture slots);
    masks, logits and memory sizes must equal the plain graph path frame by frame -- a mis-ordered stream would hand the
    decoder another frame's features."""
    from cutie_b200.config import default_config
    from cutie_b200.inference.inference_core import InferenceCore
    from cutie_b200.model.cutie import CUTIE
    from oracle.synth import synthetic_state_dict, synthetic_video
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = default_config(mem_every=3, max_mem_frames=3)
    net = CUTIE(cfg).eval()
    net.load_state_dict(synthetic_state_dict(net.state_dict(), 0))
    net = net.cuda()
    if optimised:
        net.optimize_for_inference()
    a, b = InferenceCore(net, cfg=cfg, use_cuda_graphs=True), InferenceCore(net, cfg=cfg, use_cuda_graphs=True)
    T = 16
    frames, mask = synthetic_video(T + 1, 240, 432, 3, seed=9)
    fd = frames.cuda()
    worst = 0.0
    with torch.inference_mode():
        for ti in range(T):
            kw = dict(objects=[1, 2, 3]) if ti == 0 else {}
            args = (fd[ti], mask.cuda()) if ti == 0 else (fd[ti],)
            # every third call hands over a DIFFERENT tensor than announced: the look-ahead must be discarded, not used
            announced = fd[ti + 1] if ti % 5 != 4 else fd[ti + 1].clone()
            pa = a.step(*args, next_image=announced, **kw)
            pb = b.step(*args, **kw)
            worst = max(worst, float((pa - pb).abs().max()))
            assert worst < 2e-2, (ti, worst)
            assert a.memory.work_mem.size(0) == b.memory.work_mem.size(0)
            assert (a.output_prob_to_mask(pa) != b.output_prob_to_mask(pb)).float().mean() < 1e-3
    print('look-ahead vs plain: max |prob diff| =', worst)
    assert len(a._graphs._enc) == 2 and len(b._graphs._enc) == 1      # the second capture slot exists only with look-ahead


@pytest.mark.parametrize('cin,cout,hw', [(256, 128, (60, 108)), (128, 128, (120, 216))])
def test_decoder_resblock_channels_last_variant_matches_nchw(cin, cout, hw):
    """ObjResBlock (the decoder's residual blocks) with channels-last weight twins vs the NCHW form, at the 480p shapes."""
    from cutie_b200.model.blocks import ObjResBlock, fold, unfold
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    blk = ObjResBlock(cin, cout).cuda().eval()
    g = torch.randn(1, 3, cin, *hw, device='cuda')
    with torch.inference_mode():
        ref = blk(g)
        blk.make_channels_last_twins()
        x = fold(g).contiguous(memory_format=torch.channels_last)
        got = unfold(blk._forward4(x, *blk.cl_twins), 1)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 2e-4 * float(ref.abs().max())


@pytest.mark.parametrize('shape', [(3, 128, 120, 216), (2, 16, 9, 7), (1, 5, 1, 1)])
def test_conv3x3_c1_kernel_matches_cudnn(shape):
    import cutie_b200.kernels as K_
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(shape[1])
    x = torch.randn(*shape, generator=g).cuda()
    conv = torch.nn.Conv2d(shape[1], 1, 3, padding=1).cuda()
    with torch.inference_mode():
        want = conv(torch.relu(x))
        got = K_.conv3x3_c1(x, conv.weight, conv.bias, relu_input=True)
    assert got.shape == want.shape and float((got - want).abs().max()) <= 2e-5 * max(1.0, float(want.abs().max()))
