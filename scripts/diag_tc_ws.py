"""Diagnostic (GPU box): inspect the filter's workspace (counts, thresholds) after a 3-level call."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cutie_b200.kernels as K_
L = K_.lib()
g = torch.Generator().manual_seed(0)
B, N, Q, k = 1, 70001, 256, 30
key = torch.randn(B, N, 64, generator=g).cuda(); shr = (1 + torch.randn(B, N, generator=g) ** 2).cuda()
qk = torch.randn(B, 64, Q, generator=g).cuda(); qe = torch.sigmoid(torch.randn(B, 64, Q, generator=g)).cuda()
print('plan levels', K_.affinity_plan_levels(N, k))
idx = torch.empty(B, Q, 32, dtype=torch.int32, device='cuda'); w = torch.empty(B, Q, 32, device='cuda'); sim = torch.empty(B, Q, 32, device='cuda')
nb = L.cutie_affinity_workspace_bytes(B, Q, N, k)
ws = torch.zeros(nb, dtype=torch.uint8, device='cuda')
P, I = ctypes.c_void_p, ctypes.c_int64
st = L.cutie_affinity_topk(1, (P*1)(key.data_ptr()), (P*1)(shr.data_ptr()), (I*1)(N), (I*1)(key.stride(0)), (I*1)(shr.stride(0)),
    P(qk.data_ptr()), P(qe.data_ptr()), I(B), I(64), I(Q), k, 32, P(idx.data_ptr()), P(w.data_ptr()), P(sim.data_ptr()), P(0), I(N),
    P(ws.data_ptr()), ctypes.c_size_t(nb), P(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize(); print('status', st)
al = lambda x: (x + 255) // 256 * 256
o_ci = 0; o_ce = al(B*Q*4096*4); o_cnt = o_ce + al(B*Q*4096*4); o_dm = o_cnt + al(B*Q*4); o_e0 = o_dm + al(B*Q*4); o_e1 = o_e0 + al(B*Q*4)
cnt = ws[o_cnt:o_cnt+B*Q*4].view(torch.int32); dm = ws[o_dm:o_dm+B*Q*4].view(torch.float32)
e0 = ws[o_e0:o_e0+B*Q*4].view(torch.float32); e1 = ws[o_e1:o_e1+B*Q*4].view(torch.float32)
print('final-level count: mean', float(cnt.float().mean()), 'max', int(cnt.max()), 'min', int(cnt.min()))
print('dmax mean', float(dm.mean()), 'emax0 (after level0) mean', float(e0.mean()), 'emax1 mean', float(e1.mean()))
# truth
a = qe[0].sqrt(); bq = a * qk[0]
E = torch.empty(Q, N, device='cuda')
for q0 in range(0, Q, 32):
    d = a[:, q0:q0+32].t()[:, None, :] * key[0][None] - bq[:, q0:q0+32].t()[:, None, :]
    E[q0:q0+32] = (d*d).sum(-1) * shr[0][None]
print('true kth E: stride256', float(E[:, ::256].kthvalue(k, 1)[0].mean()), 'stride16', float(E[:, ::16].kthvalue(k, 1)[0].mean()), 'full', float(E.kthvalue(k, 1)[0].mean()))
print('sim[k-1]*-8 mean', float((-8*sim[0, :, k-1]).mean()))
