"""Host logic of the product (arena bookkeeping, bucket/permanence rules, consolidation schedule, folded
attention algebra, InferenceCore.step control flow) on CPU, with the CUDA kernel wrappers replaced by the
oracle-backed emulations of tests/cpu_kernels.py.  Compared against the reference's recorded outputs."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN

LT_SMALL = dict(max_mem_frames=4, min_mem_frames=2, num_prototypes=16, max_num_tokens=60, buffer_tokens=20)


def _setup(over):
    from cutie_b200.config import default_config
    from cutie_b200.inference.inference_core import InferenceCore
    from cutie_b200.model.cutie import CUTIE
    from oracle.synth import synthetic_state_dict
    cfg = default_config(**over)
    net = CUTIE(cfg).eval()
    net.load_state_dict(synthetic_state_dict(net.state_dict(), 0))
    return cfg, net, InferenceCore(net, cfg=cfg)


def _sizes(proc):
    m = proc.memory
    row = []
    for b in sorted(m.work_mem.buckets):
        row += [b, m.work_mem.size(b), m.work_mem.perm_size(b), m.long_mem.size(b) if m.use_long_term else 0]
    return row


@pytest.mark.parametrize('name,over,T,K', [
    ('fifo', dict(mem_every=2, max_mem_frames=3), 10, 3),
    ('longterm', dict(mem_every=1, use_long_term=True, long_term=LT_SMALL), 14, 2),
])
def test_free_running_matches_reference(cpu_kernels, name, over, T, K):
    from oracle.synth import synthetic_video
    g = np.load(os.path.join(GOLDEN, f'e2e_{name}.npz'))
    cfg, net, proc = _setup(over)
    frames, mask = synthetic_video(T, 96, 160, K, seed=3)
    li = 0
    with torch.inference_mode():
        for ti in range(T):
            prob = proc.step(frames[ti], mask, objects=list(range(1, K + 1))) if ti == 0 else proc.step(frames[ti])
            assert _sizes(proc) == [int(x) for x in g['sizes'][ti] if x >= 0]
            if ti > 0:
                assert float(np.abs(proc.last_logits.numpy() - g['logits'][li:li + 1]).max()) < 2e-4
                li += 1
    assert float((prob - torch.from_numpy(g['final_prob'])).abs().max()) < 1e-4


def test_second_bucket_and_delete_match_reference(cpu_kernels):
    from oracle.synth import synthetic_video
    g = np.load(os.path.join(GOLDEN, 'e2e_buckets.npz'))
    cfg, net, proc = _setup(dict(mem_every=2, max_mem_frames=3))
    frames, _ = synthetic_video(8, 96, 160, 3, seed=3)
    with torch.inference_mode():
        for ti in range(8):
            if ti == 0:
                prob = proc.step(frames[0], torch.from_numpy(g['first_mask']), objects=[1, 2])
            elif ti == 3:
                prob = proc.step(frames[3], torch.from_numpy(g['second_mask']), objects=[7])
                assert proc.memory.work_mem.buckets == {0: [1, 2], 1: [7]}
            elif ti == 6:
                proc.delete_objects([1])
                assert proc.memory.work_mem.buckets == {0: [2], 1: [7]}
                assert 1 in proc.memory.obj_v          # reference quirk: obj_v survives purge_except
                prob = proc.step(frames[6])
            else:
                prob = proc.step(frames[ti])
            if ti == 4:
                assert float(np.abs(proc.last_logits.numpy() - g['logits_f4']).max()) < 2e-4
            assert (proc.output_prob_to_mask(prob).numpy() == g['masks'][ti]).mean() > 0.999
    assert float(np.abs(proc.last_logits.numpy() - g['logits']).max()) < 2e-4


@pytest.mark.parametrize('name,ti,over', [
    ('fifo', 7, dict(mem_every=2, max_mem_frames=3)),
    ('longterm', 9, dict(mem_every=1, use_long_term=True, long_term=LT_SMALL)),
])
def test_teacher_forced_read_matches_reference(cpu_kernels, name, ti, over):
    """Load the reference's complete memory state into the arena, run MemoryManager.read once."""
    g = np.load(os.path.join(GOLDEN, f'read_tf_{name}_{ti}.npz'))
    cfg, net, proc = _setup(over)
    m = proc.memory
    objs = [int(o) for o in g['bucket_0']]
    proc.object_manager.add_new_objects(objs)
    perm = int(g['perm_end_0'])
    Tn = lambda k: torch.from_numpy(g[k])
    wk, wsr = Tn('work_k_0'), Tn('work_s_0')
    vals = {o: Tn(f'work_v_{o}') for o in objs}
    m.CK, m.CV = wk.shape[1], vals[objs[0]].shape[1]
    m.H, m.W = g['pix_feat'].shape[-2:]
    m.HW = m.H * m.W
    m.config_stale = False
    m.max_work_tokens = m.max_mem_frames * m.HW
    if m.use_long_term:
        m.min_work_tokens = m.min_mem_frames * m.HW
    # permanent part first (as_permanent='first' on an empty bucket), then the temporary part
    sel = Tn('work_e_0') if m.use_long_term else None
    m.work_mem.add(wk[:, :, :perm], {o: v[:, :, :perm] for o, v in vals.items()}, wsr[:, :, :perm],
                   selection=sel[:, :, :0] if sel is not None else None, as_permanent='first')
    m.work_mem.add(wk[:, :, perm:], {o: v[:, :, perm:] for o, v in vals.items()}, wsr[:, :, perm:],
                   selection=sel, as_permanent='no')
    if m.use_long_term:
        arena, runs = m.work_mem.temp_runs(0)
        pos = 0
        for r in runs:
            arena.view('use', r).copy_(Tn('work_use_0')[:, pos:pos + r[1]])
            arena.view('life', r).copy_(Tn('work_life_0')[:, pos:pos + r[1]])
            pos += r[1]
        if 'long_k_0' in g:
            m.long_mem.add(Tn('long_k_0'), {o: Tn(f'long_v_{o}') for o in objs}, Tn('long_s_0'), None,
                           supposed_bucket_id=0)
            la, lr = m.long_mem.temp_runs(0)
            la.view('use', lr[0]).copy_(Tn('long_use_0'))
            la.view('life', lr[0]).copy_(Tn('long_life_0'))
    for o in objs:
        m.sensory[o] = Tn(f'sensory_{o}')
        m.obj_v[o] = Tn(f'objv_{o}')
    m.engaged = True
    # state round-trips through the arena exactly
    assert torch.equal(m.work_mem.key[0], wk) and torch.equal(m.work_mem.shrinkage[0], wsr)
    assert torch.equal(m.work_mem.value[objs[0]], vals[objs[0]])
    with torch.inference_mode():
        out = m.read(Tn('pix_feat'), Tn('query_key'), Tn('selection'), Tn('last_mask'), net)
    for o in objs:
        ref = Tn(f'out_{o}')
        assert float((out[o] - ref).abs().max()) < 3e-4 * max(1.0, float(ref.abs().max()))
    if m.use_long_term:
        assert torch.allclose(m.work_mem.use_cnt[0], Tn('work_use_after_0'), atol=1e-5)


def test_arena_ring_wraps_and_grows(cpu_kernels):
    from cutie_b200.inference.memory_bank import KeyValueMemoryStore
    st = KeyValueMemoryStore(save_selection=True, save_usage=True, ring=True)
    st.set_capacity_hint(temp_tokens=25, perm_tokens=10)
    g = torch.Generator().manual_seed(0)
    ref_k, ref_v = None, None
    for it in range(12):
        k = torch.randn(1, 64, 10, generator=g)
        v = {3: torch.randn(1, 256, 10, generator=g)}
        s = torch.rand(1, 1, 10, generator=g) + 1
        e = torch.rand(1, 64, 10, generator=g)
        st.add(k, v, s, e, as_permanent="first")
        if it == 0:
            perm_k, perm_v = k, v[3]
            continue
        ref_k = k if ref_k is None else torch.cat([ref_k, k], -1)
        ref_v = v[3] if ref_v is None else torch.cat([ref_v, v[3]], -1)
        st.remove_old_memory(0, 20)
        ref_k, ref_v = ref_k[:, :, -20:], ref_v[:, :, -20:]
        assert st.size(0) == 10 + ref_k.shape[-1] and st.perm_size(0) == 10 and st.non_perm_size(0) == ref_k.shape[-1]
        assert torch.equal(st.key[0], torch.cat([perm_k, ref_k], -1))
        assert torch.equal(st.value[3], torch.cat([perm_v, ref_v], -1))
        assert len(st.segments(0)) <= 3
    # capacity hint was too small for a burst: the ring re-linearises into a bigger arena
    big = torch.randn(1, 64, 40, generator=g)
    st.add(big, {3: torch.randn(1, 256, 40, generator=g)}, torch.ones(1, 1, 40), torch.rand(1, 64, 40, generator=g))
    assert torch.equal(st.key[0][:, :, -40:], big) and torch.equal(st.key[0][:, :, 10:30], ref_k)
    # force_permanent ('all') prepends in the reference's logical order
    p2 = torch.randn(1, 64, 10, generator=g)
    st.add(p2, {3: torch.randn(1, 256, 10, generator=g)}, torch.ones(1, 1, 10), torch.rand(1, 64, 10, generator=g),
           as_permanent='all')
    assert st.perm_size(0) == 20 and torch.equal(st.key[0][:, :, :10], p2) and torch.equal(st.key[0][:, :, 10:20], perm_k)
    st.clear_non_permanent_memory()
    assert st.size(0) == 20 and st.non_perm_size(0) == 0
    st.purge_except([])
    assert not st.engaged() and st.num_objects == 0


def test_segment_without_memory_and_empty_mask(cpu_kernels):
    cfg, net, proc = _setup(dict())
    img = torch.rand(3, 96, 160)
    with torch.inference_mode():
        out = proc.step(img)                       # no memory yet: warns, returns zeros [1,H,W]
        assert out.shape == (1, 96, 160) and float(out.abs().sum()) == 0
        out = proc.step(img, torch.zeros(96, 160, dtype=torch.long), objects=[])
        assert out.shape == (1, 96, 160)
    from cutie_b200.inference.object_manager import ObjectManager
    om = ObjectManager()
    om.add_new_objects([4, 9])
    with pytest.raises(NotImplementedError):
        om.realize_dict({4: torch.zeros(1)})
    from cutie_b200.inference.memory_bank import KeyValueMemoryStore
    with pytest.raises(RuntimeError, match='I did not count usage'):
        KeyValueMemoryStore().get_usage(0)


def test_kernels_fail_loudly_without_cuda():
    """No CPU path in the product: the ctypes wrappers refuse CPU tensors."""
    import cutie_b200.kernels as K_
    with pytest.raises(K_.KernelError):
        K_.obj_summary_accumulate(torch.zeros(4), torch.zeros(4))


def test_key_image_tracks_every_written_row(cpu_kernels, monkeypatch):
    """The arena's tcgen05 operand image must be refreshed for exactly the rows written since the last read:
    appends, ring wrap-around, re-allocation (everything moves) and long-term compaction."""
    import cutie_b200.kernels as K_
    from cutie_b200.inference.memory_bank import KeyValueMemoryStore
    calls = []

    centres = set()

    def record(key_arena, shr_arena, phys_begin, n, image, mu=None):
        assert image.shape[1] * K_.KEY_IMAGE_TILE >= key_arena.shape[1] and image.shape[2] == K_.KEY_IMAGE_FLOATS
        assert mu is not None and mu.shape == (1, 64)
        centres.add(mu.data_ptr())
        calls.append((key_arena.data_ptr(), phys_begin, n))
    monkeypatch.setattr(K_, 'bank_key_image', record)
    st = KeyValueMemoryStore(save_selection=False, save_usage=False)
    st.set_capacity_hint(temp_tokens=30, perm_tokens=10)
    g = torch.Generator().manual_seed(0)

    def add(n, perm='no'):
        st.add(torch.randn(1, 64, n, generator=g), {1: torch.randn(1, 256, n, generator=g)}, torch.ones(1, 1, n), None,
               as_permanent=perm)

    def imaged_rows(arena):
        """physical rows of `arena` covered by image refreshes since the last reset"""
        rows = set()
        for ptr, s, n in calls:
            if ptr == arena.arrays['key'].data_ptr():
                rows |= set(range(s, s + n))
        return rows
    add(10, 'first')
    add(10), add(10)
    segs = st.segments(0)
    bk = st._b[0]
    assert all(s.key_image is not None for s in segs)
    # one key centre per bucket (the mean key of the permanent first frame), shared by every image and every segment
    assert len(centres) == 1 and all(s.key_mu.data_ptr() in centres for s in segs)
    assert torch.allclose(segs[0].key_mu, bk.perm.view('key', (0, 10)).mean(1))
    assert imaged_rows(bk.perm) == set(range(10)) and imaged_rows(bk.temp) == set(range(20))
    assert [s.phys_begin for s in segs] == [0, 0]
    calls.clear()
    st.segments(0)
    assert not calls, 'nothing was written: no refresh'
    # fill the ring (cap 30), evict the oldest 20, append 15: the new run starts again at physical row 0
    add(10)
    st.remove_old_memory(0, 10)
    add(15)
    segs = st.segments(0)
    assert imaged_rows(bk.temp) == set(range(20, 30)) | set(range(0, 15))
    assert [(s.phys_begin, s.n) for s in segs] == [(0, 10), (20, 10), (0, 15)]
    # a burst larger than the capacity re-linearises the arena: every kept row is imaged again
    calls.clear()
    old_ptr = bk.temp.arrays['key'].data_ptr()
    add(40)
    segs = st.segments(0)
    assert bk.temp.arrays['key'].data_ptr() != old_ptr
    assert imaged_rows(bk.temp) == set(range(65))
    assert segs[-1].key_image.shape[1] == K_.key_image_tiles(bk.temp.cap)
    assert len(centres) == 1, 'the centre is fixed for the life of the bucket'



def test_query_chain_records_the_separate_launch_plan(cpu_kernels, monkeypatch):
    """QueryTransformer._forward_chained (one cutie_qt_chain launch per block) against the separate-launch forward on
    the emulated kernels: same outputs, and the recorded op lists respect the chain's contract -- at most 16 ops, phases
    never decrease, no op reads an output of its own phase."""
    import cutie_b200.kernels as K_
    import cutie_b200.model.object_transformer as ot
    from cutie_b200.config import default_config
    from cutie_b200.model.cutie import CUTIE
    from oracle.synth import synthetic_state_dict
    cfg = default_config()
    net = CUTIE(cfg).eval()
    net.load_state_dict(synthetic_state_dict(net.state_dict(), 0))
    qt = net.object_transformer
    g = torch.Generator().manual_seed(2)
    pixel = torch.randn(1, 3, 256, 6, 10, generator=g)
    summ = torch.rand(1, 3, 1, 16, 257, generator=g) + 0.1
    chains = []
    run = K_.qt_chain_run
    monkeypatch.setattr(K_, 'qt_chain_run', lambda ch: (chains.append(ch), run(ch))[1])
    with torch.inference_mode():
        monkeypatch.setattr(ot, 'QT_CHAIN', False)
        want, aux_w = qt(pixel, summ)
        assert not chains
        monkeypatch.setattr(ot, 'QT_CHAIN', True)
        got, aux_g = qt(pixel, summ)
    assert len(chains) == 1 + len(qt.blocks)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), float((got - want).abs().max())
    for a, b in zip(aux_g['logits'], aux_w['logits']):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5)
    assert torch.equal(aux_g['fg_map'], aux_w['fg_map'])
    for ch in chains:
        assert 1 <= len(ch.ops) <= K_.QT_CHAIN_MAX_OPS
        phases = [op[4] for op in ch.ops]
        assert phases == sorted(phases)
        for name, args, kw, outs, phase in ch.ops:
            reads = [t for t in list(args) + [v for k, v in kw.items() if k != 'xhat_out'] if isinstance(t, torch.Tensor)]
            reads += [t for v in kw.values() if isinstance(v, tuple) for t in v if isinstance(t, torch.Tensor)]
            same_phase_outs = [o for op in ch.ops if op[4] == phase and op is not None for o in op[3] if o is not None]
            same_phase_outs += [op[2].get('xhat_out') for op in ch.ops if op[4] == phase and op[2].get('xhat_out') is not None]
            for r in reads:
                assert all(r.data_ptr() != o.data_ptr() for o in same_phase_outs), f'{name} reads an output of its own phase'


def test_threshold_seed_completion_keeps_k_distinct_valid_tokens():
    """memory_manager.complete_seeds: winners the ring dropped are replaced by tokens of the newest memory frame around the
    query's own position; the completed lists are valid, distinct per query, and untouched where they were valid."""
    from cutie_b200.inference.memory_manager import complete_seeds
    g = torch.Generator().manual_seed(0)
    B, Q, kpad, top_k, HW, frames = 2, 60, 32, 30, 60, 5
    n_total = frames * HW
    old = torch.stack([torch.stack([torch.randperm(n_total - HW, generator=g)[:kpad] for _ in range(Q)]) for _ in range(B)]).int()
    old[:, :, top_k:] = -1                                           # padding slots
    seeds = old.clone()
    dropped = torch.rand(B, Q, kpad, generator=g) < 0.3              # what the ring dropped since the last read
    seeds[dropped] = -1
    out = complete_seeds(seeds, top_k, n_total, HW, 10)               # a 6 x 10 feature map
    assert torch.equal(out[:, :, top_k:], torch.full_like(out[:, :, top_k:], -1))
    live = out[:, :, :top_k]
    assert bool((live >= 0).all()) and bool((live < n_total).all())
    kept = ~dropped[:, :, :top_k]
    assert torch.equal(live[kept], old[:, :, :top_k][kept])
    assert bool((live[~kept] >= n_total - HW).all())                 # replacements come from the newest frame only
    for b in range(B):
        for q in range(Q):
            assert live[b, q].unique().numel() == top_k
    # the first dropped slot of a query is replaced by the token at the query's own position in the newest frame
    first = dropped[:, :, :top_k].int().argmax(dim=2)
    has = dropped[:, :, :top_k].any(dim=2)
    own = (n_total - HW) + torch.arange(Q).view(1, Q).expand(B, Q)
    assert torch.equal(torch.gather(live, 2, first.unsqueeze(-1)).squeeze(-1)[has], own[has].int())
    raster = complete_seeds(seeds, top_k, n_total, HW)               # no width given: raster neighbours, same guarantees
    assert all(raster[b, q, :top_k].unique().numel() == top_k for b in range(B) for q in range(Q))
    # every slot dropped, on a map so small that 2-D neighbours would wrap onto each other: still k distinct tokens
    from cutie_b200.inference.memory_manager import _neighbour_offsets
    for width, frame in ((10, 60), (54, 1620), (9, 63), (0, 61)):
        offs = _neighbour_offsets(top_k, width, frame, torch.device('cpu'))
        assert offs[0] == 0 and len({int(o) % frame for o in offs}) == top_k, (width, frame)
    gone = torch.full((1, HW, kpad), -1, dtype=torch.int32)
    allnew = complete_seeds(gone, top_k, n_total, HW, 10)
    assert all(allnew[0, q, :top_k].unique().numel() == top_k for q in range(HW))
    # nothing to do / nothing possible
    assert torch.equal(complete_seeds(old, top_k, n_total, HW), old)
    assert torch.equal(complete_seeds(seeds, top_k, n_total, 8), seeds)          # fewer tokens per frame than top_k
