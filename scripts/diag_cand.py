"""Diagnostic (GPU box): expected candidate counts of the tcgen05 filter levels on the cfg2 bench bank."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, math
import bench
from cutie_b200.inference.inference_core import InferenceCore
from oracle.synth import synthetic_video
wl = bench.WORKLOADS['cfg2']; cfg = bench.make_cfg(wl); dev = torch.device('cuda')
net = bench.make_net(cfg).to(dev)
frames, mask = synthetic_video(4, wl['H'], wl['W'], wl['K'], seed=0)
proc = InferenceCore(net, cfg=cfg)
with torch.inference_mode():
    proc.step(frames[0].to(dev), mask.to(dev), objects=[1, 2, 3])
    for key, shr, vals in bench.synthetic_bank_chunks(wl):
        proc.memory.work_mem.add(key.to(dev), {o: vals[:, i].to(dev) for i, o in enumerate([1, 2, 3])}, shr.to(dev), None)
    segs = proc.memory.work_mem.segments(0, [])
    keys = torch.cat([s.key for s in segs], 1)[0]          # [N,64]
    shr = torch.cat([s.shrinkage for s in segs], 1)[0]     # [N]
    N = keys.shape[0]
    ms, pix = net.encode_image(frames[1].to(dev)[None].contiguous() if False else torch.nn.functional.pad(frames[1].to(dev), (5, 5, 0, 0))[None])
    qk, _, qe = net.transform_key(ms[0])
    qk, qe = qk.flatten(2)[0], qe.flatten(2)[0]            # [64,Q]
    Q = qk.shape[1]
    a = qe.sqrt(); bq = a * qk
    b2 = (qe * qk * qk).sum(0); vq = b2.sqrt()
    print('N', N, 'Q', Q, 'vq mean', float(vq.mean()), 'shr max', float(shr.max()), 'knorm2 max', float((keys**2).sum(1).max()))
    E = torch.empty(Q, N, device=dev)
    for q0 in range(0, Q, 64):
        d = a[:, q0:q0+64].t()[:, None, :] * keys[None] - bq[:, q0:q0+64].t()[:, None, :]     # [64q,N,64]
        E[q0:q0+64] = (d * d).sum(-1) * shr[None]
    P = (shr * (keys**2).sum(1)).sqrt(); R = shr.sqrt()
    Pt = torch.nn.functional.pad(P, (0, (-N) % 128)).view(-1, 128).max(1)[0].repeat_interleave(128)[:N]
    Rt = torch.nn.functional.pad(R, (0, (-N) % 128)).view(-1, 128).max(1)[0].repeat_interleave(128)[:N]
    delta = 2.0**-9 * (Pt[None] + Rt[None] * vq[:, None]) ** 2          # [Q,N]
    print('delta mean', float(delta.mean()), 'E 30th smallest mean', float(E.kthvalue(30, dim=1)[0].mean()))
    for stride_prev, stride in ((256, 16), (16, 1)):
        samp = E[:, ::stride_prev]
        emax = samp.kthvalue(30, dim=1)[0]                                 # [Q]
        sub = E[:, ::stride]; dl = delta[:, ::stride]
        passed = sub < (emax[:, None] + dl)
        strict = sub < emax[:, None]
        n = sub.shape[1]; nsplit = 11
        tiles = (n + 127) // 128; tps = (tiles + nsplit - 1) // nsplit
        per = torch.stack([passed[:, s*tps*128:(s+1)*tps*128].sum(1) for s in range(nsplit)], 1)   # [Q,nsplit]
        print(f'level stride {stride}: keys {n}, cand/query mean {float(passed.sum(1).float().mean()):.0f} max {int(passed.sum(1).max())} '
              f'(strict {float(strict.sum(1).float().mean()):.0f}); per split max {int(per.max())}, overflow(>512) frac {float((per > 512).float().mean()):.4f}; '
              f'split0 mean {float(per[:,0].float().mean()):.0f} others mean {float(per[:,1:].float().mean()):.0f}')
