"""cutie_conv_tc beside cuDNN's fp32 convolution (TF32 off) with cuDNN's AUTOTUNER on -- reported, not asserted (bench.py
measures the whole step).  In a file of its own that sorts LAST: the autotuner's picks persist in the process and change
which engines later library convolutions of the same shapes run (measured: the cfg-1 reference comparison, which needs both
sides' library convolutions to be bit-identical, flipped near-tied top-k members after these tests had run)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method='thread')]


@pytest.fixture(autouse=True)
def _restore_cudnn_flags():
    bench, tf32 = torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32
    yield
    torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32 = bench, tf32


@pytest.mark.parametrize('NB,Cin,Cout,H,W', [(3, 256, 256, 30, 54), (3, 512, 768, 30, 54), (3, 256, 128, 60, 108),
                                             (3, 128, 128, 120, 216), (1, 256, 64, 30, 54)])
def test_conv3x3_tc_time_beside_cudnn_fp32(NB, Cin, Cout, H, W):
    """Reported, not asserted (bench.py measures the whole step): device time per launch, L2-warm, beside cuDNN's fp32
    convolution (TF32 off, cudnn.benchmark on) of the same layer."""
    import cutie_b200.kernels as K_
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    x = torch.randn(NB, Cin, H, W, device='cuda')
    w = torch.randn(Cout, Cin, 3, 3, device='cuda') * 0.02
    b = torch.randn(Cout, device='cuda')
    img = K_.conv_weight_image(w)

    def timed(fn, n=30):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    t_tc = timed(lambda: K_.conv_tc(x, img, b, Cout, relu_out=True))
    t_lib = timed(lambda: F.conv2d(x, w, b, padding=1).relu_())
    flops = 2.0 * NB * H * W * Cout * Cin * 9
    print(f'[{NB},{Cin}->{Cout},{H}x{W}] tcgen05 3xTF32 {t_tc:.1f} us ({flops / t_tc / 1e6:.1f} TFLOP/s fp32-equivalent), '
          f'cuDNN fp32 {t_lib:.1f} us ({flops / t_lib / 1e6:.1f} TFLOP/s)')


@pytest.mark.parametrize('NB,Cin,Cout,H,W,k', [(1, 1024, 256, 30, 54, 1), (1, 256, 1024, 30, 54, 1), (1, 64, 64, 120, 216, 3),
                                               (1, 256, 256, 30, 54, 3), (1, 128, 512, 60, 108, 1)])
def test_trunk_layers_time_beside_cudnn_channels_last(NB, Cin, Cout, H, W, k):
    """Reported, not asserted: the ResNet bottleneck layers, channels-last on both sides."""
    import cutie_b200.kernels as K_
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    x = torch.randn(NB, Cin, H, W, device='cuda').contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, k, k, device='cuda') * 0.02).contiguous(memory_format=torch.channels_last)
    b = torch.randn(Cout, device='cuda')
    img = K_.conv_weight_image(w)

    def timed(fn, n=30):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    t_tc = timed(lambda: K_.conv_tc(x, img, b, Cout, ksize=k, relu_out=True))
    t_lib = timed(lambda: F.conv2d(x, w, b, padding=k // 2).relu_())
    print(f'trunk {k}x{k} [{NB},{Cin}->{Cout},{H}x{W}] channels-last: tcgen05 3xTF32 {t_tc:.1f} us, cuDNN fp32 {t_lib:.1f} us')
