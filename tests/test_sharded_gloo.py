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


def _stream_worker(rank, world, port, ret):
    """One video stream, working memory key-sharded over `world` ranks: every rank must produce the logits of the
    un-sharded run (up to the fp32 summation order of the all-reduced readout) while holding 1/world of the bank."""
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from tests import cpu_kernels as ck
        ck.install()
        from cutie_b200.config import default_config
        from cutie_b200.inference.inference_core import InferenceCore
        from cutie_b200.inference.sharded import shard_bounds
        from cutie_b200.model.cutie import CUTIE
        from oracle.synth import synthetic_state_dict, synthetic_video
        torch.set_num_threads(2)
        cfg = default_config(mem_every=2, max_mem_frames=3, chunk_size=2)      # 3 objects in chunks of 2
        net = CUTIE(cfg).eval()
        net.load_state_dict(synthetic_state_dict(net.state_dict(), 0))
        T, K = 8, 3
        frames, mask = synthetic_video(T, 96, 160, K, seed=3)                   # 6 x 10 = 60 tokens per frame
        sharded = InferenceCore(net, cfg=cfg, memory_shard_group=dist.group.WORLD)
        plain = InferenceCore(net, cfg=cfg)
        lo, hi = shard_bounds(60, world, rank)
        worst = 0.0
        ok = True
        with torch.inference_mode():
            for ti in range(T):
                args = (frames[ti], mask) if ti == 0 else (frames[ti],)
                kw = dict(objects=[1, 2, 3]) if ti == 0 else {}
                ps = sharded.step(*args, **kw)
                pp = plain.step(*args, **kw)
                worst = max(worst, float((ps - pp).abs().max()))
                if ti > 0:
                    worst = max(worst, float((sharded.last_logits - plain.last_logits).abs().max()))
                # same frames in memory, 1/world of their tokens here; the FIFO evicts whole frames on every rank
                frames_in_mem = plain.memory.work_mem.size(0) // 60
                ok = ok and sharded.memory.work_mem.size(0) == frames_in_mem * (hi - lo)
                ok = ok and sharded.memory.work_mem.perm_size(0) == (hi - lo)
        ret[rank] = (ok, worst)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_key_sharded_stream_matches_unsharded(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() + world * 13) % 2000
    mp.spawn(_stream_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        ok, worst = ret.get(r, (False, 1e9))
        assert ok, f'rank {r}: shard sizes wrong'
        assert worst < 2e-4, f'rank {r}: sharded stream deviates by {worst}'


def test_key_sharding_rejects_long_term():
    from cutie_b200.config import default_config
    from cutie_b200.inference.memory_manager import MemoryManager
    from cutie_b200.inference.object_manager import ObjectManager

    class _Group:          # never touched: the constructor must refuse before any collective
        pass
    import torch.distributed as dist_
    cfg = default_config(use_long_term=True)
    orig = (dist_.get_world_size, dist_.get_rank)
    dist_.get_world_size, dist_.get_rank = (lambda g=None: 2), (lambda g=None: 0)
    try:
        with pytest.raises(NotImplementedError):
            MemoryManager(cfg, ObjectManager(), shard_group=_Group())
    finally:
        dist_.get_world_size, dist_.get_rank = orig
