"""Key-sharded read host logic on CPU: world_size-2 and -3 gloo process groups, kernels
emulated by tests/cpu_kernels.py.  The sharded result must equal the single-rank read of the whole bank."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import ROOT


def _worker(rank, world, port, top_k, n_total, ret):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from tests import cpu_kernels as ck
        ck.install()
        import cutie_b200.kernels as K_
        from cutie_b200.inference.sharded import shard_bounds, sharded_read
        g = torch.Generator().manual_seed(0)
        B, Q, K = 2, 45, 2
        key = torch.randn(B, n_total, 64, generator=g)
        key[:, 7] = key[:, 3]                              # a cross-check of the global tie rule
        shr = 1 + torch.randn(B, n_total, generator=g) ** 2
        shr[:, 7] = shr[:, 3]
        vals = [torch.randn(B, n_total, 256, generator=g) for _ in range(K)]
        qk = torch.randn(B, 64, Q, generator=g)
        qe = torch.sigmoid(torch.randn(B, 64, Q, generator=g))
        lo, hi = shard_bounds(n_total, world, rank)
        seg = K_.BankSegment(key[:, lo:hi], shr[:, lo:hi], tuple(v[:, lo:hi] for v in vals))
        usage = torch.zeros(B, hi - lo, dtype=torch.int64)
        out, idx, w = sharded_read([seg], lo, n_total, qk, qe, top_k, usage_acc_local=usage)
        full = K_.BankSegment(key, shr, tuple(vals))
        uref = torch.zeros(B, n_total, dtype=torch.int64)
        ridx, rw, _ = K_.affinity_topk([full], qk, qe, top_k, usage_acc=uref)
        rout = K_.readout_gather(ridx, rw, [full])
        ok = bool(torch.equal(idx, ridx) and torch.allclose(w, rw, atol=1e-6) and
                  torch.allclose(out, rout, rtol=1e-4, atol=1e-5) and
                  float((usage - uref[:, lo:hi]).abs().max()) * 2.0 ** -40 < 1e-6)
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,top_k,n_total', [(2, 30, 200), (3, 30, 100), (2, 50, 60)])
def test_sharded_read_equals_single_rank(world, top_k, n_total):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() + world * 7 + top_k) % 2000
    mp.spawn(_worker, args=(world, port, top_k, n_total, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_shard_bounds_partition():
    from cutie_b200.inference.sharded import shard_bounds
    for n in (0, 1, 7, 50000):
        for w in (1, 2, 8):
            segs = [shard_bounds(n, w, r) for r in range(w)]
            assert segs[0][0] == 0 and segs[-1][1] == n
            assert all(segs[i][1] == segs[i + 1][0] for i in range(w - 1))
            assert max(e - b for b, e in segs) - min(e - b for b, e in segs) <= 1
