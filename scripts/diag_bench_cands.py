"""GPU box: candidate counts of the filter on the cfg2 bench bank AFTER real frames have been appended."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import cutie_b200.kernels as K_
from cutie_b200.inference.inference_core import InferenceCore
from oracle.synth import synthetic_video
wl = bench.WORKLOADS['cfg2']; cfg = bench.make_cfg(wl); dev = torch.device('cuda')
net = bench.make_net(cfg).to(dev)
T = 38
frames, mask = synthetic_video(T + 2, wl['H'], wl['W'], wl['K'], seed=0)
proc = InferenceCore(net, cfg=cfg, use_cuda_graphs=True)
L = K_.lib(); P, I = ctypes.c_void_p, ctypes.c_int64
with torch.inference_mode():
    proc.step(frames[0].to(dev), mask.to(dev), objects=[1, 2, 3])
    for key, shr, vals in bench.synthetic_bank_chunks(wl):
        proc.memory.work_mem.add(key.to(dev), {o: vals[:, i].to(dev) for i, o in enumerate([1, 2, 3])}, shr.to(dev), None)
    for t in range(1, T):
        proc.step(frames[t].to(dev))
    segs = proc.memory.work_mem.segments(0, [])
    img = torch.nn.functional.pad(frames[T].to(dev), (5, 5, 0, 0))[None]
    ms, pix = net.encode_image(img); qk, _, qe = net.transform_key(ms[0])
    qk, qe = qk.flatten(2).contiguous(), qe.flatten(2).contiguous()
    B, _, Q = qk.shape; N = sum(s.n for s in segs); k = 30
    print('segments', [s.n for s in segs], 'N', N)
    idx = torch.empty(B, Q, 32, dtype=torch.int32, device=dev); w = torch.empty(B, Q, 32, device=dev); sim = torch.empty(B, Q, 32, device=dev)
    nb = L.cutie_affinity_workspace_bytes(B, Q, N, k); ws = torch.zeros(nb, dtype=torch.uint8, device=dev)
    ns = len(segs); PA, IA = P * ns, I * ns
    st = L.cutie_affinity_topk(ns, PA(*[s.key.data_ptr() for s in segs]), PA(*[s.shrinkage.data_ptr() for s in segs]), IA(*[s.n for s in segs]),
        IA(*[s.key.stride(0) for s in segs]), IA(*[s.shrinkage.stride(0) for s in segs]), P(qk.data_ptr()), P(qe.data_ptr()), I(B), I(64), I(Q), k, 32,
        P(idx.data_ptr()), P(w.data_ptr()), P(sim.data_ptr()), P(0), I(N), P(ws.data_ptr()), ctypes.c_size_t(nb), P(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    al = lambda x: (x + 255) // 256 * 256
    CAP = 16384
    o_cnt = al(B*Q*CAP*4) * 2
    cnt = ws[o_cnt:o_cnt + B*Q*4].view(torch.int32)
    print('final-level candidates/query: mean %.0f median %.0f max %d  frac>4096 %.3f frac>16384 %.3f' % (float(cnt.float().mean()), float(cnt.float().median()), int(cnt.max()), float((cnt > 4096).float().mean()), float((cnt > CAP).float().mean())))
    print('E* (=-8*kth sim) mean %.3f' % float((-8 * sim[0, :, k-1]).mean()), ' winners in newest 20k tokens frac %.3f' % float((idx[0, :, :k] >= N - 20000).float().mean()))
    segs3 = proc.memory.work_mem.segments(0, [1, 2, 3])
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    for name, tcmin in (('tcgen05 plan', -1), ('exact scan', 1 << 40)):
        K_.set_tc_min_tokens(tcmin)
        for _ in range(2): K_.affinity_topk(segs3, qk, qe, 30)
        ts = []
        for _ in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); K_.affinity_topk(segs3, qk, qe, 30); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(name, 'isolated on the evolved bank: ms', ['%.3f' % t for t in ts])
    K_.set_tc_min_tokens(-1)
