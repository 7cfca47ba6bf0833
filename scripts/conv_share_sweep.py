"""cutie_conv_tc: device time vs (tile, chunk) units per CTA on the layers that have fewer output tiles than SMs
(needs a GPU).  Used to set the plan rule of cutie_conv_plan."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cutie_b200.kernels as K_                                                     # noqa: E402

torch.backends.cudnn.allow_tf32 = False


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


LAYERS = [('PixelFFN 3x3 256->256 @30x54 x3 (NCHW)', 3, 256, 256, 30, 54, 3, False),
          ('trunk layer3 3x3 256->256 @30x54', 1, 256, 256, 30, 54, 3, True),
          ('trunk layer3 1x1 1024->256 @30x54', 1, 1024, 256, 30, 54, 1, True),
          ('trunk layer3 1x1 256->1024 @30x54 + residual', 1, 256, 1024, 30, 54, 1, True),
          ('trunk layer2 3x3 128->128 @60x108', 1, 128, 128, 60, 108, 3, True),
          ('trunk layer2 1x1 512->128 @60x108', 1, 512, 128, 60, 108, 1, True),
          ('trunk layer2 1x1 128->512 @60x108 + residual', 1, 128, 512, 60, 108, 1, True),
          ('key projection 3x3 256->64 @30x54', 1, 256, 64, 30, 54, 3, True)]
for name, NB, Cin, Cout, H, W, k, cl in LAYERS:
    x = torch.randn(NB, Cin, H, W, device='cuda')
    z = torch.randn(NB, Cout, H, W, device='cuda')
    if cl:
        x, z = x.contiguous(memory_format=torch.channels_last), z.contiguous(memory_format=torch.channels_last)
    w = torch.randn(Cout, Cin, k, k, device='cuda') * 0.02
    b = torch.randn(Cout, device='cuda')
    img = K_.conv_weight_image(w)
    cnt = torch.zeros(8192, dtype=torch.int32, device='cuda')
    C = Cin // 32
    res = []
    for q in sorted({C, *range(1, min(C, 12) + 1), C // 2, C // 4 or 1}):
        import ctypes
        plan = (ctypes.c_int64 * 6)()
        K_.lib().cutie_conv_plan(ctypes.c_int64(NB), ctypes.c_int64(Cin), ctypes.c_int64(Cout), ctypes.c_int64(H), ctypes.c_int64(W),
                                 k, 1, q, plan)
        t = timed(lambda: K_.conv_tc(x, img, b, Cout, ksize=k, residual=z, relu_out=True, units_per_cta=q, counters=cnt))
        res.append(f'q={q}:{int(plan[4])}ctas:{t:.0f}us')
    print(name, '| tiles', int(plan[0]), 'chunks', C, '|', '  '.join(res))
