"""tcgen05 (TF32) candidate filter + exact re-rank: numerics of the tensor-core contraction itself, and
bit-identity of the filtered path with the exact fp32 scan."""
import pytest
import torch

from oracle import memory_math as mm
from tests.test_gpu_kernels import K_, check_topk, make_bank, segments_of  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture
def tc_everywhere(K_):
    K_.set_tc_min_tokens(256)
    yield
    K_.set_tc_min_tokens(-1)


@pytest.mark.parametrize('B,N,Q,cuts,scale', [(1, 128, 128, (), 1.0), (2, 300, 200, (77,), 1.0), (1, 1000, 130, (128, 500, 501), 3.0)])
def test_tf32_energy_matches_exact_within_bound(K_, B, N, Q, cuts, scale):
    key, shr, vals = make_bank(B, N, 0, seed=4)
    key = key * scale
    g = torch.Generator().manual_seed(9)
    qk = torch.randn(B, 64, Q, generator=g) * scale
    qe = torch.sigmoid(torch.randn(B, 64, Q, generator=g))
    segs = segments_of(K_, key, shr, vals, cuts)
    d = K_.debug_tc_energy(segs, qk.cuda(), qe.cuda()).cpu().double()                 # [B,Q,N]
    truth = -8.0 * mm.similarity_direct(key.transpose(1, 2), shr.unsqueeze(1), qk, qe).transpose(1, 2)   # [B,Q,N]
    # The MMA output folds the rigorous per-(token, query) error bound in:  d = E_tf32 - eps*shr_n*(|k_n| + sqrt(b2_q))^2
    # so it must be a LOWER bound of the exact energy, and never further than 2*eps*(...) below it.
    knorm = key.double().norm(dim=2)                                                   # [B,N]
    vq = (qe.double() * qk.double() ** 2).sum(1).sqrt()                                # [B,Q]
    s2 = shr.double()[:, None, :] * (knorm[:, None, :] + vq[:, :, None]) ** 2
    assert (d <= truth + 1e-9).all(), f'filter value above the exact energy by {float((d - truth).max()):.3e}'
    assert (truth - d <= 2.0 * 1.66e-3 * s2).all(), f'max slack ratio {float(((truth - d) / (1.65e-3 * s2)).max()):.3f}'
    # the TF32 contraction itself (bound removed) is far more accurate than its worst case
    e_tf32 = d + 1.65e-3 * s2
    assert float(((e_tf32 - truth).abs() / truth.abs().clamp_min(1e-6)).median()) < 1e-3


@pytest.mark.parametrize('B,N,Q,K,top_k,cuts', [
    (1, 333, 77, 2, 30, ()),                   # single level: every token is a candidate
    (2, 1000, 130, 2, 30, (128, 500, 501)),    # four segments, batch 2
    (1, 4099, 1620, 3, 30, (4000,)),           # 480p query count
    (1, 5000, 300, 1, 64, (100,)),             # kpad 64
    (1, 70001, 96, 1, 30, (1620, 30000)),      # 3 nested levels (strides 256 -> 16 -> 1)
])
def test_filtered_path_matches_oracle_and_exact_scan(K_, tc_everywhere, B, N, Q, K, top_k, cuts):
    assert K_.affinity_plan_levels(N, top_k) >= 1
    idx_tc, w_tc = check_topk(K_, B, N, Q, K, top_k, cuts)          # all oracle assertions on the filtered path
    K_.set_tc_min_tokens(1 << 40)                                   # same inputs through the exact scan only
    assert K_.affinity_plan_levels(N, top_k) == 0
    idx_ex, w_ex = check_topk(K_, B, N, Q, K, top_k, cuts)
    assert torch.equal(idx_tc, idx_ex), 'filtered selection differs from the exact scan'
    assert torch.equal(w_tc, w_ex), 'weights are not bit-identical'


def test_filter_keeps_duplicates_and_near_duplicates(K_, tc_everywhere):
    """Near-duplicate frames (the hard case of SURVEY.md Appendix B): many tokens within TF32 noise of each other."""
    g = torch.Generator().manual_seed(2)
    base = torch.randn(1, 500, 64, generator=g) * 4
    key = torch.cat([base + 1e-3 * torch.randn(1, 500, 64, generator=g) for _ in range(8)], 1)   # 4000 tokens
    shr = 1 + torch.randn(1, 4000, generator=g) ** 2
    vals = [torch.randn(1, 4000, 256, generator=g)]
    qk = base[:, :200].transpose(1, 2).contiguous() + 1e-3 * torch.randn(1, 64, 200, generator=g)
    qe = torch.sigmoid(torch.randn(1, 64, 200, generator=g))
    segs = segments_of(K_, key, shr, vals, (1500,))
    idx, w, sim = K_.affinity_topk(segs, qk.cuda(), qe.cuda(), 30, want_sim=True)
    K_.set_tc_min_tokens(1 << 40)
    idx2, w2, sim2 = K_.affinity_topk(segs, qk.cuda(), qe.cuda(), 30, want_sim=True)
    assert torch.equal(idx, idx2) and torch.equal(w, w2) and torch.equal(sim, sim2)
