"""TEST / CHECKER INFRASTRUCTURE (see oracle/memory_math.py header): used by tests/ and by bench.py's in-run parity check.

Teacher forcing: copy the CPU oracle's complete recurrent state (working / long-term memory, usage
counters, sensory memory, object summaries, last mask, frame clocks) into a product InferenceCore, so
that a single step can be compared without the chaotic amplification of a free-running recurrent net
(SURVEY.md section 7 'hard parts': parity must be measured teacher-forced per frame)."""
import torch


def load_state_from_oracle(proc, oc, device):
    from cutie_b200.inference.memory_manager import MemoryManager
    from cutie_b200.inference.object_manager import ObjectManager
    proc.object_manager = ObjectManager()
    if oc.objects:
        proc.object_manager.add_new_objects(list(oc.objects))
    m = MemoryManager(cfg=proc.cfg, object_manager=proc.object_manager)
    proc.memory = m
    proc.curr_ti, proc.last_mem_ti = oc.curr_ti, oc.last_mem_ti
    proc.last_mask = oc.last_mask.to(device) if oc.last_mask is not None else None
    if not oc.work.buckets:
        return
    d = lambda t: t.to(device).contiguous()
    any_v = next(iter(oc.work.v.values()))
    m.CK, m.CV = next(iter(oc.work.k.values())).shape[1], any_v.shape[1]
    some_s = next(iter(oc.sensory.values()))
    m.H, m.W = some_s.shape[-2:]
    m.HW = m.H * m.W
    m.config_stale = False
    m.max_work_tokens = m.max_mem_frames * m.HW
    if m.use_long_term:
        m.min_work_tokens = m.min_mem_frames * m.HW
        m.long_mem.set_capacity_hint(temp_tokens=m.max_long_tokens + m.num_prototypes)
    m.work_mem.set_capacity_hint(temp_tokens=m.max_work_tokens + m.HW, perm_tokens=m.HW)
    m.work_mem.global_bucket_id = oc.work.next_bucket
    for b, objs in oc.work.buckets.items():
        p = oc.work.perm_end.get(b, 0)
        k, s = oc.work.k[b], oc.work.s[b]
        vals = {o: oc.work.v[o] for o in objs}
        # recreate bucket ids faithfully: buckets are created in increasing id order
        m.work_mem.global_bucket_id = b
        sel = oc.work.e.get(b) if m.use_long_term else None
        if m.use_long_term and sel is None:
            sel = torch.zeros(k.shape[0], k.shape[1], 0)
        if p > 0:
            m.work_mem.add(d(k[:, :, :p]), {o: d(v[:, :, :p]) for o, v in vals.items()}, d(s[:, :, :p]),
                           selection=d(sel[:, :, :0]) if sel is not None else None, as_permanent='first')
        if k.shape[-1] > p:
            m.work_mem.add(d(k[:, :, p:]), {o: d(v[:, :, p:]) for o, v in vals.items()}, d(s[:, :, p:]),
                           selection=d(sel) if sel is not None else None, as_permanent='no')
            if m.use_long_term:
                arena, runs = m.work_mem.temp_runs(b)
                pos = 0
                for r in runs:
                    arena.view('use', r).copy_(d(oc.work.use[b][:, pos:pos + r[1]]))
                    arena.view('life', r).copy_(d(oc.work.life[b][:, pos:pos + r[1]]))
                    pos += r[1]
        if m.use_long_term and b in oc.long.buckets:
            m.long_mem.add(d(oc.long.k[b]), {o: d(oc.long.v[o]) for o in objs}, d(oc.long.s[b]), None,
                           supposed_bucket_id=b)
            if m.long_mem.save_usage:
                la, lr = m.long_mem.temp_runs(b)
                la.view('use', lr[0]).copy_(d(oc.long.use[b]))
                la.view('life', lr[0]).copy_(d(oc.long.life[b]))
    m.work_mem.global_bucket_id = oc.work.next_bucket
    for o, t in oc.sensory.items():
        m.sensory[o] = d(t)
    for o, t in oc.obj_v.items():
        m.obj_v[o] = d(t).clone()
    m.engaged = oc.engaged


def export_state_to_oracle(proc, oc):
    """The reverse direction: the product's live recurrent state (arena-resident memory bank included) into a fresh
    OracleCore, channel-major on the CPU as the reference keeps it -- so that the oracle can re-compute ONE frame from
    exactly the state the CUDA path is in (bench.py's `parity_check` after the timed region).  FIFO mode only."""
    m = proc.memory
    assert not m.use_long_term, 'export_state_to_oracle covers FIFO working memory (the bench workloads)'
    def cpu(t):
        c = t.detach().float().cpu()
        return c.clone() if c.data_ptr() == t.data_ptr() else c      # never alias the product's state (CPU dry runs)
    oc.objects = [int(o) for o in proc.object_manager.all_obj_ids]
    oc.curr_ti, oc.last_mem_ti = proc.curr_ti, proc.last_mem_ti
    oc.last_mask = cpu(proc.last_mask) if proc.last_mask is not None else None
    oc.engaged = bool(m.engaged)
    oc.HW = m.HW
    w = m.work_mem
    oc.work.next_bucket = w.global_bucket_id
    keys, shrs = w.key, w.shrinkage                     # channel-major exports of the arena (one per bucket)
    for b, objs in w.buckets.items():
        oc.work.buckets[b] = [int(o) for o in objs]
        oc.work.k[b], oc.work.s[b] = cpu(keys[b]), cpu(shrs[b])
        oc.work.perm_end[b] = int(w.perm_size(b))
    del keys, shrs
    for o, b in w._objs.items():                        # one object at a time: a cfg2 value export is 423 MB
        oc.work.v[int(o)] = cpu(w._export(b, ('val', o), True))
    oc.sensory = {int(o): cpu(t) for o, t in m.sensory.items()}
    oc.obj_v = {int(o): cpu(t) for o, t in m.obj_v.items()}
    return oc


class SelectionReconciler:
    """Near-tie arbitration for teacher-forced comparisons.

    The CUDA path ranks by the cancellation-free fp32 form, the oracle (like the reference) by the fp32
    three-term expansion whose rounding noise (~1e-4 absolute on O(100) terms) exceeds the gap between the
    k-th and (k+1)-th similarity on a few queries per frame.  For every query where the two top-k SETS
    differ, this hook checks -- against the float64 direct-form ground truth -- that the CUDA selection is a
    valid top-k (every chosen token is within `rel_tol` of the true k-th value) and, if so, lets the oracle
    adopt it, so everything downstream can be compared to 1e-3.  An invalid selection raises."""

    def __init__(self, top_k, rel_tol=2e-5, max_frac=0.02):
        self.top_k, self.rel_tol, self.max_frac = top_k, rel_tol, max_frac
        self.gpu_idx = None        # [B,Q,kpad] int32 captured from kernels.affinity_topk
        self.flips = 0
        self.queries = 0

    def __call__(self, bucket, mk, ms, qk, qe, sim, idx):
        from oracle import memory_math as mm
        k = self.top_k
        g = self.gpu_idx[:, :, :k].transpose(1, 2).long().cpu()            # [B,k,Q]
        diff = (g.sort(1)[0] != idx.sort(1)[0]).any(1)                      # [B,Q]
        self.queries += diff.numel()
        if not diff.any():
            return idx
        out = idx.clone()
        for b, q in diff.nonzero().tolist():
            truth = mm.similarity_direct(mk[b:b + 1], ms[b:b + 1], qk[b:b + 1, :, q:q + 1], qe[b:b + 1, :, q:q + 1])[0, :, 0]
            kth = torch.topk(truth, k)[0][-1]
            chosen = truth[g[b, :, q]]
            tol = self.rel_tol * float(kth.abs()) + 1e-7
            assert float((kth - chosen).max()) <= tol, \
                f'CUDA top-k picked a token {float((kth - chosen).max()):.3e} below the true k-th (tol {tol:.1e})'
            out[b, :, q] = g[b, :, q]
            self.flips += 1
        assert self.flips <= self.max_frac * self.queries + 2, 'too many near-tie disagreements'
        return out


class ForegroundReconciler:
    """Near-tie arbitration for the foreground test of _get_aux_mask (object_transformer.py:185-192), the network's other
    discrete decision: fg[k,p] = (logit_k >= max over {bg, 1..K}).  On a pixel whose two largest log-odds differ by less
    than fp32 rounding of the mask_pred accumulation, the CUDA kernel and the oracle can legitimately disagree, and with
    random-init weights ONE flipped pixel moves the next logits by 4e-2 (measured on the bike clip).  For every pixel
    where the two maps differ this hook checks that the margin is below `tol` (in log-odds, from the ORACLE's own aux
    logits) and, if so, lets the oracle adopt the CUDA map; a disagreement on a pixel with a real margin raises.

    `gpu_fg`: list (one per aux stage, in call order) of uint8/bool [B,K,HW] maps captured from kernels.qt_aux_mask."""

    def __init__(self, tol=2e-4):
        self.tol = tol
        self.gpu_fg = []
        self.flips = 0
        self.pixels = 0

    def __call__(self, stage, aux_logits, fg):
        from oracle.transformer import aggregate_logits
        if stage >= len(self.gpu_fg):
            return fg
        g = self.gpu_fg[stage].to(torch.bool).cpu().reshape(fg.shape)
        diff = g != fg
        self.pixels += fg.numel()
        if not diff.any():
            return fg
        lo = aggregate_logits(aux_logits.sigmoid(), dim=1).flatten(2)            # [B, 1+K, HW]
        top2 = lo.topk(2, dim=1)[0]
        margin = (top2[:, 0] - top2[:, 1]).abs()                                  # [B, HW]
        bad = diff.any(1) & (margin > self.tol)
        assert not bad.any(), (f'foreground maps differ on {int(bad.sum())} pixel(s) with a margin up to '
                               f'{float(margin[bad].max()):.3e} (> {self.tol:.0e}) at aux stage {stage}')
        self.flips += int(diff.sum())
        return g
