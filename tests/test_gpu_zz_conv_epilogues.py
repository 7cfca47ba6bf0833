"""conv + bias (+ residual) + ReLU through cuDNN's fused graph (cutie_b200/model/fuse.ConvEpilogueFuser) against the
three-launch form, on the GPU.  These are PyTorch/cuDNN stages either side of the hot path (kept as library calls);
what is asserted is that switching the call form changes nothing beyond fp32 rounding, whatever each layer's
on-device trial decided.  (File name sorts last on purpose: it exercises cuDNN engines, not cutie_b200 kernels.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu


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
        out_f = net.pixel_encoder(img)                      # trials run here, then the winners
        out_f2 = net.pixel_encoder(img)
        fz.enabled = False
        out_u = net.pixel_encoder(img)
        fz.enabled = True
    rep = fz.report()
    print('conv epilogues (pixel encoder, 240p, fp32):', rep)
    assert rep['fused'] + rep['three_launch'] >= 40        # ResNet-50 stages 1-3: 1 stem + 39 convs with a ReLU
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
    assert not off.conv_epilogues.decisions
