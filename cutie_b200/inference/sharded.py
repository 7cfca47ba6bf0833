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

Long-term memory under key sharding (memory_manager.py:283-358, kv_memory_store.py:209-242) adds three exchanges, all
on memory frames only:

    prototype selection   global top-P of usage = top-P of the all-gathered local top-P lists           (select_top)
    potentiation          every shard runs the dense softmax over ITS candidates against all P prototypes
                          (cutie_consolidate_partial) and returns its affinity maximum and exp-sum per prototype;
                          NCCL all_gather of those per-shard maxima / sums -> each shard's exact softmax share ->
                          all_reduce(sum) of the re-weighted partial prototypes                         (combine_partial_softmax)
    obsolete removal      global top-max_size of the long-term usage, survivors re-dealt to the ranks in contiguous
                          blocks of the global ranking (balanced, and identical for every batch entry)   (fetch_rows)
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


# ---- long-term memory under key sharding ---------------------------------------------------------------------------

def select_top(values_local: torch.Tensor, k: int, group=None):
    """Global top-k of a per-token score sharded over the group.  values_local [B, n_local] (n_local may differ between
    ranks, may be 0).  Returns (src_rank [B,k], src_idx [B,k]) int64, identical on every rank, in descending order of the
    score; ties resolve to the lower rank, then to the local top-k order.  One all_gather of the local top-k lists
    (the global top-k is a subset of their union)."""
    world = dist.get_world_size(group)
    B, n_local = values_local.shape
    dev = values_local.device
    k_loc = min(k, n_local)
    v = torch.full((B, k), float('-inf'), dtype=torch.float32, device=dev)
    i = torch.full((B, k), -1, dtype=torch.int64, device=dev)
    if k_loc > 0:
        tv, ti = torch.topk(values_local, k=k_loc, dim=1, sorted=True)
        v[:, :k_loc], i[:, :k_loc] = tv, ti
    all_v = torch.empty(world * B, k, dtype=torch.float32, device=dev)
    all_i = torch.empty(world * B, k, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_v, v, group=group)
    dist.all_gather_into_tensor(all_i, i, group=group)
    flat_v = all_v.view(world, B, k).permute(1, 0, 2).reshape(B, world * k)
    flat_i = all_i.view(world, B, k).permute(1, 0, 2).reshape(B, world * k)
    order = torch.sort(flat_v, dim=1, descending=True, stable=True)[1][:, :k]
    src_idx = flat_i.gather(1, order)
    if bool((src_idx < 0).any()):
        raise ValueError(f'fewer than {k} tokens over all shards')
    return order // k, src_idx


def fetch_rows(rows_local: Sequence[torch.Tensor], src_rank: torch.Tensor, src_idx: torch.Tensor, group=None) -> torch.Tensor:
    """out[b, j, :] = the row src_idx[b, j] of rank src_rank[b, j]'s concatenated `rows_local` ([B, n_i, C] token-major
    runs), on every rank: each rank gathers the rows it owns into a zero tensor, all_reduce(sum) -- exact, every row has
    one non-zero contributor."""
    rank = dist.get_rank(group)
    B, m = src_rank.shape
    mine = src_rank == rank
    C = rows_local[0].shape[2] if rows_local else None
    if C is None:
        raise ValueError('fetch_rows needs the row width: pass at least one (possibly empty) run')
    dev = src_rank.device
    out = torch.zeros(B, m, C, dtype=torch.float32, device=dev)
    runs = [r for r in rows_local if r.shape[1] > 0]
    if runs and bool(mine.any()):
        got = torch.empty(B, m, C, dtype=torch.float32, device=dev)
        K_.bank_gather(runs, torch.where(mine, src_idx, torch.zeros_like(src_idx)).contiguous(), got)
        out = torch.where(mine.unsqueeze(-1), got, out)
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def combine_partial_softmax(partial: torch.Tensor, local_max: torch.Tensor, local_sumexp: torch.Tensor, group=None) -> torch.Tensor:
    """partial [B, P, C]: softmax-weighted sums over THIS shard's candidates, normalised by the shard's own statistics
    local_max / local_sumexp [B, P] (cutie_consolidate_partial).  Returns the sums under the softmax over ALL shards'
    candidates (memory_utils.py:68-71 evaluated shard-wise): all_gather of the per-shard affinity maxima and exp-sums,
    share_r = sumexp_r exp(max_r - M) / sum_r' sumexp_r' exp(max_r' - M), then all_reduce(sum) of share_r * partial_r."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B, P = local_max.shape
    stats = torch.stack([local_max, local_sumexp], dim=-1).contiguous()
    all_stats = torch.empty(world * B, P, 2, dtype=torch.float32, device=stats.device)
    dist.all_gather_into_tensor(all_stats, stats, group=group)
    all_stats = all_stats.view(world, B, P, 2)
    mx, se = all_stats[..., 0], all_stats[..., 1]
    big = mx.max(dim=0, keepdim=True)[0]
    share = se * torch.exp(mx - big)
    share = share / share.sum(dim=0, keepdim=True)
    out = (partial * share[rank].unsqueeze(-1)).contiguous()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out
