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


# ---------------------------------------------------------------------------------------------------------
# key image (bulk-copy producer of the stride-1 filter level)
# ---------------------------------------------------------------------------------------------------------
def _image_offsets(row, elem):
    """Byte offset of operand element `elem` (0..135) of token row `row` inside a 69632-byte tile
    (csrc/tc_operand.cuh: 4 SWIZZLE_128B K-blocks of 32 tf32 + one un-swizzled 8-element tail block)."""
    if elem < 128:
        blk, chunk, within = elem >> 5, (elem & 31) >> 2, elem & 3
        return blk * 16384 + row * 128 + ((chunk ^ (row & 7)) << 4) + within * 4
    e = elem - 128
    return 4 * 16384 + (e >> 2) * 2048 + (row >> 3) * 128 + (row & 7) * 16 + (e & 3) * 4


def test_key_image_layout_and_values(K_):
    g = torch.Generator().manual_seed(3)
    B, cap = 2, 1000
    key = (torch.randn(B, cap, 64, generator=g) * 2).cuda()
    shr = (1 + torch.randn(B, cap, generator=g) ** 2).cuda()
    tiles = K_.key_image_tiles(cap)
    img = torch.full((B, tiles, K_.KEY_IMAGE_FLOATS), 7.0, device='cuda')
    p0, n = 77, 600                                   # unaligned range: rows outside it must stay untouched
    K_.bank_key_image(key, shr, p0, n, img)
    img, key, shr = img.cpu(), key.cpu(), shr.cpu()
    flat = img.reshape(B, tiles, -1)
    rows = torch.arange(cap)
    t, r = rows // 128, rows % 128
    eps = 1.65e-3
    for b in range(B):
        inside = (rows >= p0) & (rows < p0 + n)
        ln = shr[b][:, None] * key[b]                                         # shr k    (fp32, same op order)
        sq = ln * key[b]                                                      # shr k^2
        for c in (0, 1, 31, 32, 63):
            for which, want in ((0, sq[:, c]), (64, ln[:, c])):
                off = torch.tensor([_image_offsets(int(x), which + c) // 4 for x in r])
                got = flat[b, t, off]
                assert torch.equal(got[inside], want[inside]), f'element {which + c}'
                assert (got[~inside] == 7.0).all(), 'rows outside the range were written'
        P = (shr[b] * key[b].pow(2).sum(1)).sqrt() * 1.002
        R = shr[b].sqrt() * 1.002
        tail = [shr[b], torch.zeros(cap), shr[b], -eps * P * P, -2 * eps * P * R, -eps * R * R, torch.zeros(cap), torch.zeros(cap)]
        for e, want in enumerate(tail):
            off = torch.tensor([_image_offsets(int(x), 128 + e) // 4 for x in r])
            got = flat[b, t, off]
            torch.testing.assert_close(got[inside], want[inside], rtol=2e-5, atol=1e-9)
            # the bound factors must never be smaller than the exact ones (they are rounded UP by 1.002)
        offP = torch.tensor([_image_offsets(int(x), 128 + 3) // 4 for x in r])
        exactP2 = eps * (shr[b].double() * key[b].double().pow(2).sum(1))
        assert (-(flat[b, t, offP][inside]).double() >= exactP2[inside]).all()


def _arena_bank(K_, B, layout, seed):
    """layout: list of (capacity, phys_begin, n) -- one arena per entry, the segment is rows [phys, phys+n)."""
    g = torch.Generator().manual_seed(seed)
    segs, keys, shrs = [], [], []
    for cap, p0, n in layout:
        key = (torch.randn(B, cap, 64, generator=g) * 1.5).cuda()
        shr = (1 + torch.randn(B, cap, generator=g) ** 2).cuda()
        img = torch.full((B, K_.key_image_tiles(cap), K_.KEY_IMAGE_FLOATS), float('nan'), device='cuda')
        K_.bank_key_image(key, shr, p0, n, img)        # everything outside the segment stays NaN on purpose
        segs.append(K_.BankSegment(key[:, p0:p0 + n], shr[:, p0:p0 + n], (), img, p0))
        keys.append(key[:, p0:p0 + n]), shrs.append(shr[:, p0:p0 + n])
    return segs, torch.cat(keys, 1), torch.cat(shrs, 1)


@pytest.mark.parametrize('B,Q,top_k,layout', [
    (1, 300, 30, [(9000, 0, 9000)]),                                        # aligned, one arena
    (1, 260, 30, [(3000, 1, 2999), (20000, 12345, 7000), (20000, 0, 5001)]),    # ring wrap: tail run + head run
    (2, 130, 30, [(700, 130, 500), (8000, 127, 7000), (8000, 7999, 1), (6000, 128, 3000)]),   # 4 runs, batch 2
    (1, 96, 64, [(80000, 3, 70001)]),                                       # 3 levels, kpad 64
])
def test_image_path_is_bit_identical_to_exact_scan(K_, tc_everywhere, B, Q, top_k, layout):
    segs, key, shr = _arena_bank(K_, B, layout, seed=11)
    N = key.shape[1]
    assert K_.affinity_plan_levels(N, top_k) >= 2
    g = torch.Generator().manual_seed(5)
    qk = (torch.randn(B, 64, Q, generator=g) * 1.5).cuda()
    qe = torch.sigmoid(torch.randn(B, 64, Q, generator=g)).cuda()
    before = K_.image_level_launches()
    acc = torch.zeros(B, N, dtype=torch.int64, device='cuda')
    idx, w, sim = K_.affinity_topk(segs, qk, qe, top_k, usage_acc=acc, want_sim=True)
    assert K_.image_level_launches() == before + 1, 'the stride-1 level did not use the key image'
    plain = [K_.BankSegment(s.key, s.shrinkage, ()) for s in segs]          # same bank, in-kernel producers
    idx_p, w_p, sim_p = K_.affinity_topk(plain, qk, qe, top_k, want_sim=True)
    assert K_.image_level_launches() == before + 1
    K_.set_tc_min_tokens(1 << 40)                                           # exact fp32 scan only
    acc_x = torch.zeros(B, N, dtype=torch.int64, device='cuda')
    idx_x, w_x, sim_x = K_.affinity_topk(plain, qk, qe, top_k, usage_acc=acc_x, want_sim=True)
    for a, b_ in ((idx, idx_x), (w, w_x), (sim, sim_x), (idx_p, idx_x), (w_p, w_x), (acc, acc_x)):
        assert torch.equal(a, b_)
    # and against float64 truth: the selected set is the true top-k up to fp32 near-ties
    truth = mm.similarity_direct(key.cpu().transpose(1, 2), shr.cpu().unsqueeze(1), qk.cpu(), qe.cpu(), dtype=torch.float64)
    n_dec, n_dec_eq, _, _ = mm.topk_set_agreement(idx[:, :, :top_k].cpu().long().transpose(1, 2), truth, top_k,
                                                  rel_noise=1e-5)
    assert n_dec > 0 and n_dec_eq == n_dec
