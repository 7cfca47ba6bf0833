"""Key-sharded memory read for ONE video stream across G GPUs.

Each rank owns a slice of the memory bank (keys, shrinkage, values of every object) and the full query.
One exchange step:

    local   cutie_affinity_topk over the rank's tokens           -> k candidates (similarity, local index) per query
    NCCL    all_gather of [B, Q, kpad] similarities + global indices (8*kpad*Q bytes per rank: 0.39 MB @480p)
    local   cutie_topk_merge over the G lists                      -> global top-k, softmax weights (identical on all ranks)
    local   cutie_readout_gather over the winners this rank owns   -> partial readout [B, K, CV, Q]
    NCCL    all_reduce(sum) of the partial readouts                (K*CV*Q*4 bytes: 5 MB @ 480p / 3 objects)

Usage counters stay with the owning shard.  The result equals the single-GPU read of the concatenated bank
(ties resolve to the lower GLOBAL index because candidates carry global indices).  The reference has no
multi-GPU inference path; this is new functionality named by BASELINE.json's north_star.
"""
from typing import Optional, Sequence

import torch
import torch.distributed as dist

from cutie_b200 import kernels as K_
from cutie_b200.kernels import BankSegment


def shard_bounds(n_total: int, world: int, rank: int):
    """Contiguous, balanced partition of [0, n_total): returns (begin, end) of `rank`'s slice."""
    base, rem = divmod(n_total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def sharded_topk(local_segments: Sequence[BankSegment], index_offset: int, n_total: int, qk: torch.Tensor,
                 qe: torch.Tensor, top_k: int, group=None, usage_acc_local: Optional[torch.Tensor] = None,
                 marks: Optional[list] = None):
    """The exchange step: local top-k -> all_gather of candidates -> merge.  local_segments: this rank's tokens
    (global indices index_offset .. index_offset + n_local; only keys/shrinkage are read).
    Returns (idx_local, w_local, idx, w): the global winners [B,Q,kpad] (identical on every rank) and the same lists
    restricted to the tokens this rank owns (local indices, -1 / weight 0 elsewhere) for sharded_gather.
    `marks` (bench only): a list that receives CUDA events after the local top-k and after the all-gather."""
    world = dist.get_world_size(group)
    n_local = sum(s.n for s in local_segments)
    B, CK, Q = qk.shape
    kpad = K_.kpad_for(top_k)
    dev = qk.device
    k_local = min(top_k, n_local)
    if k_local > 0:
        idx_l, _, sim_l = K_.affinity_topk(local_segments, qk, qe, k_local, want_sim=True)
        if idx_l.shape[-1] != kpad:                      # fewer than 33 local tokens but top_k > 32
            pad = kpad - idx_l.shape[-1]
            idx_l = torch.nn.functional.pad(idx_l, (0, pad), value=-1)
            sim_l = torch.nn.functional.pad(sim_l, (0, pad), value=0.0)
        gidx = torch.where(idx_l >= 0, idx_l + index_offset, idx_l)
    else:
        gidx = torch.full((B, Q, kpad), -1, dtype=torch.int32, device=dev)
        sim_l = torch.zeros(B, Q, kpad, device=dev)
    if marks is not None:
        marks.append(torch.cuda.Event(enable_timing=True)); marks[-1].record()
    all_sim = torch.empty(world * B, Q, kpad, device=dev)
    all_idx = torch.empty(world * B, Q, kpad, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(all_sim, sim_l.contiguous(), group=group)
    dist.all_gather_into_tensor(all_idx, gidx.contiguous(), group=group)
    if marks is not None:
        marks.append(torch.cuda.Event(enable_timing=True)); marks[-1].record()
    part_val = all_sim.view(world, B, Q, kpad).permute(1, 0, 2, 3).contiguous()
    part_idx = all_idx.view(world, B, Q, kpad).permute(1, 0, 2, 3).contiguous()
    usage_global = None
    if usage_acc_local is not None:
        usage_global = torch.zeros(B, n_total, dtype=torch.int64, device=dev)
    idx, w, _ = K_.topk_merge(part_val, part_idx, top_k, n_total, usage_acc=usage_global)
    if usage_acc_local is not None:
        usage_acc_local.add_(usage_global[:, index_offset:index_offset + n_local])
    mine = (idx >= index_offset) & (idx < index_offset + n_local)
    idx_local = torch.where(mine, idx - index_offset, torch.full_like(idx, -1))
    w_local = torch.where(mine, w, torch.zeros_like(w))
    return idx_local, w_local, idx, w


def sharded_gather(idx_local: torch.Tensor, w_local: torch.Tensor, local_segments: Sequence[BankSegment], group=None):
    """Partial readout of the winners this rank owns, summed over the ranks: [B,K,CV,Q], identical everywhere."""
    if not local_segments or sum(s.n for s in local_segments) == 0 or len(local_segments[0].values) == 0:
        raise ValueError('every rank needs at least one token and one object value array')
    out = K_.readout_gather(idx_local, w_local, local_segments)
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def sharded_read(local_segments: Sequence[BankSegment], index_offset: int, n_total: int, qk: torch.Tensor,
                 qe: torch.Tensor, top_k: int, group=None, usage_acc_local: Optional[torch.Tensor] = None):
    """local_segments: this rank's tokens (global indices index_offset .. index_offset + n_local).
    Returns (readout [B,K,CV,Q] identical on every rank, idx [B,Q,kpad] global, weights [B,Q,kpad])."""
    idx_local, w_local, idx, w = sharded_topk(local_segments, index_offset, n_total, qk, qe, top_k, group,
                                              usage_acc_local)
    return sharded_gather(idx_local, w_local, local_segments, group), idx, w
