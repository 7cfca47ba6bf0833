"""The memory read at the shapes of BASELINE.json's other configurations: parity-test cases,
not bench lines.

  cfg 3  720p, 5 objects, long-term memory: HW = 3600 queries, bank = long (10 000) | permanent (3 600) | working ring in
         two pieces -> 4 segments, N = 46 000; usage counters on (long-term mode)
  cfg 5  1080p, 10 objects, 50 000 keys sharded over 8 GPUs: HW = 8 160 queries against one shard's 6 250 keys (the local
         top-k each rank contributes), and against the unsharded 50 000-key bank

The kernels run on ALL queries; the float64 oracle judges a random subset of them (the queries are independent), the
size-independent properties are checked on all: k winners per query, weights sum to 1, usage mass == number of queries,
sorted winners."""
import pytest
import torch

from oracle import memory_math as mm
from tests.test_gpu_kernels import K_, make_bank, segments_of      # noqa: F401  (K_ is a fixture)

# first run of these cases is at round end: never let one of them wedge the suite (thread method: the process exits)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method='thread')]


def _check(K_, N, Q, K, cuts, top_k=30, n_probe=96, seed=0):
    B = 1
    key, shr, vals = make_bank(B, N, K, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    qk = torch.randn(B, 64, Q, generator=g)
    qe = torch.sigmoid(torch.randn(B, 64, Q, generator=g))
    segs = segments_of(K_, key, shr, vals, cuts)
    usage = torch.zeros(B, N, dtype=torch.int64).cuda()
    idx, w, sim = K_.affinity_topk(segs, qk.cuda(), qe.cuda(), top_k, usage_acc=usage, want_sim=True)
    out = K_.readout_gather(idx, w, segs)
    torch.cuda.synchronize()
    idx, w, sim, out, usage = idx.cpu(), w.cpu(), sim.cpu(), out.cpu(), usage.cpu()
    # ---- properties on every query ----
    kk = idx[:, :, :top_k].long()
    assert (idx[:, :, top_k:] == -1).all() and (w[:, :, top_k:] == 0).all()
    assert (kk >= 0).all() and (kk < N).all()
    assert (kk.sort(-1)[0][:, :, 1:] != kk.sort(-1)[0][:, :, :-1]).all(), 'duplicate winners'
    assert torch.allclose(w.sum(-1), torch.ones(B, Q), atol=1e-5)
    assert (sim[:, :, :top_k - 1] >= sim[:, :, 1:top_k]).all()
    assert abs(float(usage.double().sum()) * 2.0 ** -40 - B * Q) < 1e-3 * B * Q
    assert out.shape == (B, K, 256, Q) and torch.isfinite(out).all()
    # ---- float64 oracle on a subset of the queries ----
    probe = torch.randperm(Q, generator=g)[:n_probe].sort()[0]
    truth = mm.similarity_direct(key.transpose(1, 2), shr.unsqueeze(1), qk[:, :, probe], qe[:, :, probe],
                                 dtype=torch.float64)                       # [B,N,n_probe]
    kp = kk[:, probe].transpose(1, 2)                                       # [B,k,n_probe]
    got_sim = sim[:, probe, :top_k].transpose(1, 2).double()
    ref_sim = torch.gather(truth, 1, kp)
    assert torch.allclose(got_sim, ref_sim, rtol=2e-5, atol=1e-5)
    n_dec, n_dec_eq, _, n_all = mm.topk_set_agreement(kp, truth, top_k, 4e-6)
    assert n_dec_eq == n_dec, f'{n_dec - n_dec_eq} decidable queries picked a different top-{top_k} set'
    assert n_dec >= 0.9 * n_all
    ref_w = torch.softmax(ref_sim, dim=1).float()
    assert torch.allclose(w[:, probe, :top_k].transpose(1, 2), ref_w, rtol=1e-4, atol=1e-6)
    aff = mm.scatter_affinity(kp, w[:, probe, :top_k].transpose(1, 2).contiguous(), N)
    vstack = torch.stack([v.transpose(1, 2) for v in vals], 1)              # [B,K,CV,N]
    assert torch.allclose(out[:, :, :, probe], mm.readout(aff, vstack), rtol=1e-4, atol=2e-5)


def test_cfg3_720p_5_objects_long_term_bank(K_):
    _check(K_, N=46000, Q=3600, K=5, cuts=(10000, 13600, 40000))


def test_cfg5_1080p_10_objects_one_key_shard(K_):
    _check(K_, N=6250, Q=8160, K=10, cuts=(680,))


def test_cfg5_1080p_10_objects_unsharded_bank(K_):
    _check(K_, N=50000, Q=8160, K=10, cuts=(8160,), n_probe=64)
