"""Where one steady-state frame of the bench workload spends its device time (no ncu: kineto/CUPTI kernel records of eager
steps, so durations are warm-cache, in-stream ones), grouped (a) by kernel name, (b) by convolution shape.

    python scripts/step_kernel_times.py [--steps 5] > profiles/rNN_step_kernel_times.md       (needs a GPU)
"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--workload', default='cfg2')
    a = ap.parse_args()
    import bench
    import cutie_b200.kernels as K_
    from cutie_b200.inference.inference_core import InferenceCore
    from cutie_b200.utils.synth import synthetic_video
    from torch.profiler import ProfilerActivity, profile
    K_.lib()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    wl = bench.WORKLOADS[a.workload]
    dev = torch.device('cuda:0')
    cfg = bench.make_cfg(wl)
    net = bench.make_net(cfg).to(dev)
    net.optimize_for_inference()
    warm = 6
    frames, mask = synthetic_video(warm + a.steps + 2, wl['H'], wl['W'], wl['K'], seed=0)
    objs = list(range(1, wl['K'] + 1))
    proc = InferenceCore(net, cfg=cfg, use_cuda_graphs=False)
    with torch.inference_mode():
        proc.step(frames[0].to(dev), mask.to(dev), objects=objs)
        for key, shr, vals in bench.synthetic_bank_chunks(wl):
            proc.memory.work_mem.add(key.to(dev), {o: vals[:, i].to(dev) for i, o in enumerate(objs)}, shr.to(dev), None,
                                     as_permanent='no')
        fd = frames.to(dev)
        for t in range(1, warm + 1):
            proc.step(fd[t])
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
            for t in range(warm + 1, warm + 1 + a.steps):
                proc.step(fd[t])
            torch.cuda.synchronize()
    n = a.steps
    kern = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            k = kern.setdefault(ev.name[:110], [0, 0.0])
            k[0] += 1
            k[1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
    total = sum(v[1] for v in kern.values())
    print(f'# Device time of one steady-state frame ({a.workload}, eager, fp32 / TF32 off, cudnn.benchmark on), mean of {n} frames '
          f'(one of them a memory frame)\n')
    print(f'total kernel time per frame: {total / n / 1e3:.3f} ms in {sum(v[0] for v in kern.values()) / n:.0f} launches\n')
    print('| kernel | launches / frame | us / frame | us / launch |\n|---|---:|---:|---:|')
    for name, (c, t) in sorted(kern.items(), key=lambda kv: -kv[1][1])[:70]:
        print(f'| `{name}` | {c / n:.1f} | {t / n:.1f} | {t / c:.1f} |')
    ours = sum(t for name, (c, t) in kern.items() if 'cutie::' in name)
    qt = sum(t for name, (c, t) in kern.items() if 'cutie::qt_' in name)
    print(f'\nkernels of this repo: {ours / n:.1f} us / frame; of which object-transformer (qt_*): {qt / n:.1f} us / frame\n')
    print('## Convolutions by shape (aten::cudnn_convolution / aten::conv2d inputs)\n')
    print('| op | input shapes | calls / frame | device us / frame | us / call |\n|---|---|---:|---:|---:|')
    rows = []
    for e in prof.key_averages(group_by_input_shape=True):
        if 'conv' in e.key and 'cudnn' in e.key or e.key in ('aten::convolution_relu', 'aten::cudnn_convolution_relu',
                                                             'aten::cudnn_convolution_add_relu'):
            dt = e.device_time_total if hasattr(e, 'device_time_total') else e.cuda_time_total
            if dt > 0:
                rows.append((dt, e.key, str(e.input_shapes)[:110], e.count))
    for dt, key, shp, cnt in sorted(rows, reverse=True)[:50]:
        print(f'| {key} | `{shp}` | {cnt / n:.1f} | {dt / n:.1f} | {dt / cnt:.1f} |')
    print(f'\nconvolution ops total: {sum(r[0] for r in rows) / n / 1e3:.3f} ms / frame')


if __name__ == '__main__':
    main()
