"""Key-sharded read over NCCL on >= 2 GPUs (skipped on a single-GPU box): sharded == single-GPU result."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n_total, Q, K, ret):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        import cutie_b200.kernels as K_
        from cutie_b200.inference.sharded import shard_bounds, sharded_read
        g = torch.Generator().manual_seed(0)
        B, top_k = 1, 30
        key = torch.randn(B, n_total, 64, generator=g).cuda()
        shr = (1 + torch.randn(B, n_total, generator=g) ** 2).cuda()
        vals = [torch.randn(B, n_total, 256, generator=g).cuda() for _ in range(K)]
        qk = torch.randn(B, 64, Q, generator=g).cuda()
        qe = torch.sigmoid(torch.randn(B, 64, Q, generator=g)).cuda()
        lo, hi = shard_bounds(n_total, world, rank)
        seg = K_.BankSegment(key[:, lo:hi], shr[:, lo:hi], tuple(v[:, lo:hi] for v in vals))
        usage = torch.zeros(B, hi - lo, dtype=torch.int64, device='cuda')
        out, idx, w = sharded_read([seg], lo, n_total, qk, qe, top_k, usage_acc_local=usage)
        full = K_.BankSegment(key, shr, tuple(vals))
        uref = torch.zeros(B, n_total, dtype=torch.int64, device='cuda')
        ridx, rw, _ = K_.affinity_topk([full], qk, qe, top_k, usage_acc=uref)
        rout = K_.readout_gather(ridx, rw, [full])
        torch.cuda.synchronize()
        ret[rank] = bool(torch.equal(idx, ridx) and torch.equal(w, rw) and
                         torch.allclose(out, rout, rtol=1e-5, atol=1e-5) and torch.equal(usage, uref[:, lo:hi]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
@pytest.mark.parametrize('n_total,Q,K', [(5000, 1620, 3), (50000, 1620, 3),
                                         (50000, 8160, 10)])      # cfg 5: 1080p queries, 10 objects, 50k keys
def test_sharded_read_nccl(n_total, Q, K):
    world = min(torch.cuda.device_count(), 8)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29641 + (n_total + Q) % 97, n_total, Q, K, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def _stream_worker(rank, world, port, ret):
    """One stream with its working memory key-sharded over the ranks (MemoryManager(shard_group=...)) against the
    un-sharded run on the same GPU: same logits up to the summation order of the all-reduced readout."""
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        from cutie_b200.config import default_config
        from cutie_b200.inference.inference_core import InferenceCore
        from cutie_b200.model.cutie import CUTIE
        from oracle.synth import synthetic_state_dict, synthetic_video
        cfg = default_config(mem_every=2, max_mem_frames=4)
        net = CUTIE(cfg).eval()
        net.load_state_dict(synthetic_state_dict(net.state_dict(), 0))
        net = net.cuda()
        T, K = 4, 3            # free-running: keep the clip short (random weights amplify rounding differences)
        frames, mask = synthetic_video(T, 240, 432, K, seed=3)
        sharded = InferenceCore(net, cfg=cfg, memory_shard_group=dist.group.WORLD)
        plain = InferenceCore(net, cfg=cfg)
        worst = 0.0
        with torch.inference_mode():
            for ti in range(T):
                args = (frames[ti].cuda(), mask.cuda()) if ti == 0 else (frames[ti].cuda(),)
                kw = dict(objects=[1, 2, 3]) if ti == 0 else {}
                sharded.step(*args, **kw)
                plain.step(*args, **kw)
                if ti > 0:
                    worst = max(worst, float((sharded.last_logits - plain.last_logits).abs().max()))
        torch.cuda.synchronize()
        ret[rank] = worst
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
def test_key_sharded_stream_nccl():
    world = min(torch.cuda.device_count(), 8)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_stream_worker, args=(world, 29877, ret), nprocs=world, join=True)
    assert all(ret.get(r, 1e9) < 1e-3 for r in range(world)), dict(ret)


def _long_term_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from tests import sharded_memory_case
        out = []
        for B in (1, 2):                       # 2 = the flip-augmentation batch
            out.append(sharded_memory_case.run(torch.device('cuda', rank), h=15, w=27, K=3, steps=12, B=B, P=32))
        torch.cuda.synchronize()
        ret[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
def test_key_sharded_long_term_memory_nccl():
    """Long-term memory under key sharding with the real kernels over NCCL (tests/sharded_memory_case.py): global prototype
    ranking, cutie_consolidate_partial + all-gathered affinity maxima / exp-sums, obsolete-feature removal with re-dealt
    survivors -- reads equal the un-sharded MemoryManager's at every step."""
    world = min(torch.cuda.device_count(), 8)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_long_term_worker, args=(world, 29911, ret), nprocs=world, join=True)
    for r in range(world):
        for ok, worst, trace in ret.get(r, [(False, 1e9, [])]):
            assert ok, f'rank {r}: store sizes wrong (long-term trace {trace})'
            assert max(trace) == 128 and any(a > b for a, b in zip(trace, trace[1:])), f'no obsolete-feature removal in {trace}'
            assert worst < 1e-4, f'rank {r}: sharded reads deviate by {worst}'
