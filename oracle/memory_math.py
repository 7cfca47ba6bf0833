"""CPU oracle for the pixel-memory readout math.

TEST INFRASTRUCTURE ONLY.  Nothing in cutie_b200/ imports this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may.  It restates, in
plain torch CPU ops, what the reference computes in
  cutie/model/utils/memory_utils.py:7-46   (get_similarity)
  cutie/model/utils/memory_utils.py:49-77  (do_softmax)
  cutie/inference/memory_manager.py:77-88  (MemoryManager._readout)
  cutie/inference/memory_manager.py:329-358 (consolidation)
and is pinned against fixtures produced by the unmodified reference (tests/golden/make_golden.py).
"""
import math
from typing import Optional, Tuple

import torch


def similarity_expanded(mk: torch.Tensor, ms: Optional[torch.Tensor], qk: torch.Tensor,
                        qe: Optional[torch.Tensor]) -> torch.Tensor:
    """memory_utils.py:28-44 -- the three-term expansion the reference evaluates.

    mk [B,CK,N] memory keys, ms [B,1,N] shrinkage (or None), qk [B,CK,Q] query keys,
    qe [B,CK,Q] query selection (or None).  Returns S [B,N,Q].
    """
    B, CK, N = mk.shape
    mkt = mk.transpose(1, 2)  # [B,N,CK]
    if qe is not None:
        term_aa = torch.matmul(mkt * mkt, qe)               # sum_c mk^2 qe
        term_ab = 2.0 * torch.matmul(mkt, qk * qe)          # 2 sum_c mk qk qe
        term_bb = (qe * qk * qk).sum(dim=1, keepdim=True)   # sum_c qe qk^2   [B,1,Q]
        s = term_ab - term_aa - term_bb
    else:
        term_aa = (mk * mk).sum(dim=1).unsqueeze(2)         # [B,N,1]
        s = 2.0 * torch.matmul(mkt, qk) - term_aa
    scale = 1.0 / math.sqrt(CK)
    if ms is not None:
        s = s * ms.reshape(B, N, 1) * scale                 # memory_utils.py:42
    else:
        s = s * scale
    return s


def similarity_direct(mk: torch.Tensor, ms: torch.Tensor, qk: torch.Tensor, qe: torch.Tensor,
                      dtype=torch.float64, chunk: int = 4096) -> torch.Tensor:
    """Cancellation-free form S[n,q] = -ms[n]/sqrt(CK) * sum_c qe[c,q] (mk[c,n]-qk[c,q])^2.

    Algebraically equal to similarity_expanded; evaluated in `dtype`
    (float64 by default) it is the ground truth used to judge near-tie top-k selections.
    """
    B, CK, N = mk.shape
    Q = qk.shape[-1]
    mk, ms, qk, qe = (t.to(dtype) for t in (mk, ms, qk, qe))
    out = torch.empty(B, N, Q, dtype=dtype)
    for n0 in range(0, N, chunk):
        d = mk[:, :, n0:n0 + chunk, None] - qk[:, :, None, :]          # [B,CK,n,Q]
        out[:, n0:n0 + chunk] = -(qe[:, :, None, :] * d * d).sum(dim=1)
    return out * ms.reshape(B, N, 1) / math.sqrt(CK)


def topk_softmax(sim: torch.Tensor, top_k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """memory_utils.py:57-61.  Returns (indices [B,k,Q] int64, weights [B,k,Q]).

    The reference exponentiates the k winners WITHOUT subtracting the maximum (:60) and divides by
    their sum (:61); reproduced literally here (NaN if every winner underflows, like the reference).
    """
    values, indices = torch.topk(sim, k=top_k, dim=1)
    e = values.exp()
    return indices, e / e.sum(dim=1, keepdim=True)


def scatter_affinity(indices: torch.Tensor, weights: torch.Tensor, N: int) -> torch.Tensor:
    """memory_utils.py:63-66: densify to [B,N,Q] with exactly k non-zeros per query column."""
    B, k, Q = indices.shape
    return torch.zeros(B, N, Q, dtype=weights.dtype).scatter_(1, indices, weights)


def usage_from_affinity(affinity: torch.Tensor) -> torch.Tensor:
    """memory_utils.py:74-75: usage[b,n] = sum_q A[b,n,q]."""
    return affinity.sum(dim=2)


def dense_softmax(sim: torch.Tensor) -> torch.Tensor:
    """memory_utils.py:68-71 (top_k=None branch: max-subtracted softmax over the memory axis)."""
    m = sim.max(dim=1, keepdim=True)[0]
    e = (sim - m).exp()
    return e / e.sum(dim=1, keepdim=True)


def readout(affinity: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """memory_manager.py:77-88.  v [B,C,N] -> [B,C,Q];  v [B,K,C,N] -> [B,K,C,Q]."""
    if v.dim() == 3:
        return torch.matmul(v, affinity)
    B, K, C, N = v.shape
    return torch.matmul(v.reshape(B, K * C, N), affinity).reshape(B, K, C, -1)


def sparse_readout(indices: torch.Tensor, weights: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """The same contraction as readout(scatter_affinity(...), v) evaluated over the k winners only
    (what the CUDA gather kernel does).  v [B,K,C,N] -> [B,K,C,Q]."""
    B, K, C, N = v.shape
    k, Q = indices.shape[1:]
    out = torch.zeros(B, K, C, Q, dtype=v.dtype)
    for b in range(B):
        g = v[b][:, :, indices[b].reshape(-1)].reshape(K, C, k, Q)
        out[b] = (g * weights[b][None, None]).sum(dim=2)
    return out


def consolidate(cand_key: torch.Tensor, cand_shrinkage: torch.Tensor, cand_selection: torch.Tensor,
                cand_values: dict, usage: torch.Tensor, num_prototypes: int):
    """memory_manager.py:329-358: prototype selection + potentiation.

    cand_key [B,CK,Nc], cand_shrinkage [B,1,Nc], cand_selection [B,CK,Nc], cand_values {obj: [B,CV,Nc]},
    usage [B,Nc] (normalised use_cnt/life_cnt).  Returns (proto_key [B,CK,P], {obj: [B,CV,P]},
    proto_shrinkage [B,1,P], proto_indices [B,P]).
    """
    B = cand_key.shape[0]
    pk, pe, pidx = [], [], []
    for b in range(B):
        _, idx = torch.topk(usage[b], k=num_prototypes, dim=-1, sorted=True)   # :339
        idx = idx.flatten()
        pidx.append(idx)
        pk.append(cand_key[b][:, idx])
        pe.append(cand_selection[b][:, idx])
    pk, pe = torch.stack(pk, 0), torch.stack(pe, 0)
    sim = similarity_expanded(cand_key, cand_shrinkage, pk, pe)                 # :348-349
    aff = dense_softmax(sim)                                                    # :350
    pv = {k: readout(aff, v) for k, v in cand_values.items()}                   # :353
    ps = readout(aff, cand_shrinkage)                                           # :356
    return pk, pv, ps, torch.stack(pidx, 0)


def topk_set_agreement(sim_test_idx: torch.Tensor, truth64: torch.Tensor, top_k: int,
                       rel_noise: float):
    """Compare a [B,k,Q] index selection with the float64 ground-truth similarity `truth64` [B,N,Q].

    A query column counts as *decidable* when the gap between the k-th and (k+1)-th true similarity
    exceeds rel_noise * |k-th similarity| -- the rounding noise of an fp32 evaluation of that element
    (SURVEY.md Appendix B take-away 5).  Returns
    (n_decidable, n_decidable_equal, n_total_equal, n_total).
    """
    B, N, Q = truth64.shape
    kk = min(top_k + 1, N)
    tv, ti = torch.topk(truth64, k=kk, dim=1)
    if kk > top_k:
        gap = tv[:, top_k - 1] - tv[:, top_k]
        decidable = gap > rel_noise * tv[:, top_k - 1].abs()
    else:
        decidable = torch.ones(B, Q, dtype=torch.bool)
    true_sets = ti[:, :top_k].sort(dim=1)[0]
    test_sets = sim_test_idx.sort(dim=1)[0]
    same = (true_sets == test_sets).all(dim=1)          # [B,Q]
    return int(decidable.sum()), int((same & decidable).sum()), int(same.sum()), int(same.numel())
