"""Shared body of the key-sharded LONG-TERM memory case (tests/test_sharded_gloo.py on CPU with emulated kernels,
tests/test_gpu_sharded.py over NCCL with the real ones): two MemoryManagers -- one key-sharded over the process group, one
holding everything -- are fed the same synthetic memory frames and queries.  No network in the loop, so nothing amplifies
rounding: the only recurrence is through the usage counters (bit-deterministic) into the prototype ranking, and the read
results must agree to fp32 summation order at every step while consolidations and an obsolete-feature removal happen."""
import torch
import torch.distributed as dist


def run(device, h=6, w=10, K=2, steps=12, B=1, P=8):
    from cutie_b200.config import default_config
    from cutie_b200.inference.memory_manager import MemoryManager
    from cutie_b200.inference.object_manager import ObjectManager
    from cutie_b200.inference.sharded import shard_bounds
    world, rank = dist.get_world_size(), dist.get_rank()
    HW = h * w
    cfg = default_config(use_long_term=True, top_k=30,
                         long_term=dict(max_mem_frames=4, min_mem_frames=2, num_prototypes=P, max_num_tokens=5 * P,
                                        buffer_tokens=P + P // 4))
    objs = list(range(1, K + 1))
    sharded = MemoryManager(cfg, ObjectManager(), shard_group=dist.group.WORLD)
    plain = MemoryManager(cfg, ObjectManager())
    g = torch.Generator().manual_seed(11)
    base = torch.randn(B, 64, h, w, generator=g)
    worst, trace, ok = 0.0, [], True
    for t in range(steps):
        # temporally coherent keys (a slow drift around a base frame), as a video produces them
        key = (base + 0.3 * torch.randn(B, 64, h, w, generator=g)).to(device)
        shr = (1 + torch.randn(B, 1, h, w, generator=g) ** 2).to(device)
        sel = torch.sigmoid(torch.randn(B, 64, h, w, generator=g)).to(device)
        val = torch.randn(B, K, 256, h, w, generator=g).to(device)
        for m in (sharded, plain):
            m.add_memory(key, shr, val, None, objs, selection=sel, as_permanent='first' if t == 0 else 'no')
        for _ in range(2):                                  # two reads per memory frame: usage accumulates between consolidations
            qk = (base + 0.3 * torch.randn(B, 64, h, w, generator=g)).to(device)
            qe = torch.sigmoid(torch.randn(B, 64, h, w, generator=g)).to(device)
            a = sharded.read_visual(qk, qe, objs)
            b = plain.read_visual(qk, qe, objs)
            worst = max(worst, float((a - b).abs().max()))
        n_long = plain.long_mem.size(0) if plain.long_mem.engaged(0) else 0
        mine = sharded.long_mem.size(0) if sharded.long_mem.engaged(0) else 0
        counts = sharded._long_counts.get(0, [0] * world)
        lo, hi = shard_bounds(HW, world, rank)
        ok = ok and sum(counts) == n_long and counts[rank] == mine
        ok = ok and sharded.work_mem.size(0) * HW == plain.work_mem.size(0) * (hi - lo)
        trace.append(n_long)
    return ok, worst, trace
