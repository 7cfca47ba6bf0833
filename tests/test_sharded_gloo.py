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


def _long_term_stream_worker(rank, world, port, ret):
    """The same comparison with LONG-TERM memory on (memory_manager.py:283-358): consolidations (global prototype ranking,
    shard-wise potentiation combined through the all-gathered affinity maxima / exp-sums) and one obsolete-feature removal
    (global usage ranking, survivors re-dealt) happen inside the clip; the sharded stream must track the un-sharded one and
    hold the expected share of both stores."""
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
        P = 8
        cfg = default_config(mem_every=1, use_long_term=True,
                             long_term=dict(max_mem_frames=4, min_mem_frames=2, num_prototypes=P, max_num_tokens=40,
                                            buffer_tokens=10))
        net = CUTIE(cfg).eval()
        net.load_state_dict(synthetic_state_dict(net.state_dict(), 0))
        T, K = 13, 2
        frames, mask = synthetic_video(T, 96, 160, K, seed=5)                   # 60 tokens per frame
        sharded = InferenceCore(net, cfg=cfg, memory_shard_group=dist.group.WORLD)
        plain = InferenceCore(net, cfg=cfg)
        lo, hi = shard_bounds(60, world, rank)
        worst, ok, long_trace = 0.0, True, []
        with torch.inference_mode():
            for ti in range(T):
                args = (frames[ti], mask) if ti == 0 else (frames[ti],)
                kw = dict(objects=[1, 2]) if ti == 0 else {}
                ps = sharded.step(*args, **kw)
                pp = plain.step(*args, **kw)
                worst = max(worst, float((ps - pp).abs().max()))
                if ti > 0:
                    worst = max(worst, float((sharded.last_logits - plain.last_logits).abs().max()))
                ok = ok and sharded.memory.work_mem.size(0) * 60 == plain.memory.work_mem.size(0) * (hi - lo)
                n_long = plain.memory.long_mem.size(0) if plain.memory.long_mem.engaged(0) else 0
                mine = sharded.memory.long_mem.size(0) if sharded.memory.long_mem.engaged(0) else 0
                counts = sharded.memory._long_counts.get(0, [0] * world)
                ok = ok and sum(counts) == n_long and counts[rank] == mine
                long_trace.append(n_long)
        ret[rank] = (ok, worst, long_trace)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_key_sharded_long_term_stream_matches_unsharded(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 33500 + (os.getpid() + world * 17) % 2000
    mp.spawn(_long_term_stream_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        ok, worst, trace = ret.get(r, (False, 1e9, []))
        assert ok, f'rank {r}: shard sizes wrong (long-term trace {trace})'
        assert max(trace) == 32 and 22 + 8 in trace, f'the clip must consolidate and remove obsolete features: {trace}'
        assert worst < 2e-4, f'rank {r}: sharded long-term stream deviates by {worst}'


def _memory_case_worker(rank, world, port, B, ret):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from tests import cpu_kernels as ck
        ck.install()
        from tests import sharded_memory_case
        torch.set_num_threads(2)
        ret[rank] = sharded_memory_case.run('cpu', B=B)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,B', [(2, 1), (3, 2)])
def test_key_sharded_long_term_memory_level(world, B):
    """MemoryManager-level (no network): reads of the key-sharded long-term + working memory equal the un-sharded reads
    at every step; B = 2 is the flip-augmentation batch (per-entry rankings, same counts on every rank)."""
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 37500 + (os.getpid() + world * 23) % 2000
    mp.spawn(_memory_case_worker, args=(world, port, B, ret), nprocs=world, join=True)
    for r in range(world):
        ok, worst, trace = ret.get(r, (False, 1e9, []))
        assert ok, f'rank {r}: store sizes wrong (long-term trace {trace})'
        assert max(trace) == 32 and any(a > b for a, b in zip(trace, trace[1:])), f'no obsolete-feature removal in {trace}'
        assert worst < 1e-4, f'rank {r}: sharded reads deviate by {worst}'


def _select_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from tests import cpu_kernels as ck
        ck.install()
        from cutie_b200.inference import sharded as S
        g = torch.Generator().manual_seed(1)
        B, n, C, k = 2, 37, 5, 9
        score = torch.rand(B, n, generator=g)
        score[:, 30] = score[:, 2]                                        # a tie across shards: lower rank first
        rows = torch.randn(B, n, C, generator=g)
        sizes = [0, 30, 7] if world == 3 else [30, 7]                      # ragged shards, one of them empty
        lo = sum(sizes[:rank])
        sl = slice(lo, lo + sizes[rank])
        src_rank, src_idx = S.select_top(score[:, sl], k, None)
        want_v, want_i = torch.sort(score, dim=1, descending=True, stable=True)
        gidx = src_idx + torch.tensor([sum(sizes[:r]) for r in range(world)])[src_rank]
        ok = torch.equal(gidx, want_i[:, :k])
        got = S.fetch_rows([rows[:, sl][:, :3], rows[:, sl][:, 3:]], src_rank, src_idx, None)
        ok = ok and torch.equal(got, torch.stack([rows[b][want_i[b, :k]] for b in range(B)]))
        # shard-wise softmax sums against the dense softmax over everything
        sim = torch.randn(B, n, 4, generator=g) * 5
        val = torch.randn(B, n, 6, generator=g)
        dense = torch.einsum('bnp,bnc->bpc', torch.softmax(sim, dim=1), val)
        if sizes[rank]:
            mx = sim[:, sl].max(dim=1)[0]
            e = (sim[:, sl] - mx.unsqueeze(1)).exp()
            se = e.sum(dim=1)
            part = torch.einsum('bnp,bnc->bpc', e, val[:, sl]) / se.unsqueeze(-1)
        else:
            mx, se, part = torch.full((B, 4), float('-inf')), torch.zeros(B, 4), torch.zeros(B, 4, 6)
        full = S.combine_partial_softmax(part, mx, se, None)
        ok = ok and torch.allclose(full, dense, rtol=1e-5, atol=1e-6)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_selection_fetch_and_softmax_combination(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 35500 + (os.getpid() + world * 19) % 2000
    mp.spawn(_select_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_key_sharding_needs_a_prototype_per_rank():
    from cutie_b200.config import default_config
    from cutie_b200.inference.memory_manager import MemoryManager
    from cutie_b200.inference.object_manager import ObjectManager

    class _Group:          # never touched: the constructor must refuse before any collective
        pass
    import torch.distributed as dist_
    cfg = default_config(use_long_term=True, long_term=dict(num_prototypes=1))
    orig = (dist_.get_world_size, dist_.get_rank)
    dist_.get_world_size, dist_.get_rank = (lambda g=None: 2), (lambda g=None: 0)
    try:
        with pytest.raises(ValueError):
            MemoryManager(cfg, ObjectManager(), shard_group=_Group())
    finally:
        dist_.get_world_size, dist_.get_rank = orig
