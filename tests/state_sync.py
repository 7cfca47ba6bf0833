"""Teacher forcing: copy the CPU oracle's complete recurrent state (working / long-term memory, usage
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
