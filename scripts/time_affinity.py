"""GPU box: time cutie_affinity_topk alone on the bench bank (exact scan vs tcgen05 filter plan)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import cutie_b200.kernels as K_
from cutie_b200.inference.inference_core import InferenceCore
from oracle.synth import synthetic_video
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'cfg2']; cfg = bench.make_cfg(wl); dev = torch.device('cuda')
net = bench.make_net(cfg).to(dev)
frames, mask = synthetic_video(3, wl['H'], wl['W'], wl['K'], seed=0)
proc = InferenceCore(net, cfg=cfg)
with torch.inference_mode():
    proc.step(frames[0].to(dev), mask.to(dev), objects=[1, 2, 3])
    for key, shr, vals in bench.synthetic_bank_chunks(wl):
        proc.memory.work_mem.add(key.to(dev), {o: vals[:, i].to(dev) for i, o in enumerate([1, 2, 3])}, shr.to(dev), None)
    segs = proc.memory.work_mem.segments(0, [1, 2, 3])
    img = torch.nn.functional.pad(frames[1].to(dev), (5, 5, 0, 0))[None]
    ms, pix = net.encode_image(img); qk, _, qe = net.transform_key(ms[0])
    qk, qe = qk.flatten(2).contiguous(), qe.flatten(2).contiguous()
    N = sum(s.n for s in segs)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    for name, tcmin in (('tcgen05 plan', -1), ('exact scan', 1 << 40)):
        K_.set_tc_min_tokens(tcmin)
        print(name, 'levels', K_.affinity_plan_levels(N, 30))
        for _ in range(3): K_.affinity_topk(segs, qk, qe, 30)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); idx, w, _ = K_.affinity_topk(segs, qk, qe, 30); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print('  affinity_topk ms: min %.3f median %.3f' % (min(ts), sorted(ts)[5]))
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = K_.readout_gather(idx, w, segs); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print('  readout_gather ms: min %.3f median %.3f' % (min(ts), sorted(ts)[5]))
    K_.set_tc_min_tokens(-1)
