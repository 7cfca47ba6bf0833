"""CPU oracle for the whole per-frame loop: a compact, dense restatement of the reference's
InferenceCore.step + MemoryManager + KeyValueMemoryStore on plain torch CPU ops.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/memory_math.py header).  Used (a) as the checker for the
CUDA path on the GPU box at sizes beyond the committed fixtures, (b) as the timed CPU arm of bench.py
(`cpu_baseline`, `--impl reference`).  It follows, step for step:
  cutie/inference/inference_core.py:172-328   (step), :123-170 (_segment), :71-121 (_add_memory)
  cutie/inference/memory_manager.py:112-208   (read), :210-296 (add_memory), :309-358 (consolidation)
  cutie/inference/kv_memory_store.py:55-149   (add), :151-162, :164-242 (usage / sieve / obsolete removal)
with the same data layout the reference uses (channel-major tensors grown by torch.cat, dense
[N,HW] affinity, dense readout GEMM).  The convolutional stages are the product's PyTorch modules run
on CPU (pinned to the reference's by the free-running fixtures of tests/test_oracle_golden.py: full-frame
logits of the unmodified reference within 2e-4); the memory
math and the object transformer are oracle/memory_math.py and oracle/transformer.py.
It is pinned by the committed reference fixtures (tests/test_oracle_golden.py).
"""
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from oracle import memory_math as mm
from oracle.transformer import aggregate_logits, query_transformer


def _pad16(x):
    h, w = x.shape[-2:]
    nh, nw = -(-h // 16) * 16, -(-w // 16) * 16
    lh, lw = (nh - h) // 2, (nw - w) // 2
    pad = (lw, nw - w - lw, lh, nh - h - lh)
    return F.pad(x, pad), pad


class _Store:
    """kv_memory_store.py, channel-major, torch.cat growth."""

    def __init__(self, save_selection, save_usage):
        self.save_selection, self.save_usage = save_selection, save_usage
        self.next_bucket = 0
        self.buckets: Dict[int, List[int]] = {}
        self.k, self.s, self.e, self.v = {}, {}, {}, {}
        self.use, self.life = {}, {}
        self.perm_end: Dict[int, int] = {}

    def size(self, b):
        return self.k[b].shape[-1] if b in self.k else 0

    def non_perm(self, b):
        return self.size(b) - self.perm_end.get(b, 0)

    def add(self, key, values, shrinkage, selection, supposed_bucket=-1, as_permanent='no'):
        ne = key.shape[-1]
        B = key.shape[0]

        def put(d, name, new, prepend):
            if name in d:
                d[name] = torch.cat([new, d[name]] if prepend else [d[name], new], -1)
            else:
                d[name] = new
        if supposed_bucket >= 0:
            enabled = [supposed_bucket]
            for o, val in values.items():
                put(self.v, o, val, as_permanent == 'all')
            self.buckets[supposed_bucket] = list(values.keys())
        else:
            enabled, new_b = [], None
            for o, val in values.items():
                if o in self.v:
                    put(self.v, o, val, as_permanent == 'all')
                    b = [bb for bb, objs in self.buckets.items() if o in objs][0]
                else:
                    self.v[o] = val
                    if new_b is None:
                        new_b = self.next_bucket
                        self.next_bucket += 1
                        self.buckets[new_b] = []
                    self.buckets[new_b].append(o)
                    b = new_b
                if b not in enabled:
                    enabled.append(b)
        for b in enabled:
            perm = False
            if as_permanent == 'all':
                self.perm_end[b] = self.perm_end.get(b, 0) + ne
                perm = True
            elif as_permanent == 'first' and self.perm_end.get(b, 0) == 0:
                self.perm_end[b] = ne
                perm = True
            put(self.k, b, key, perm)
            put(self.s, b, shrinkage, perm)
            if not perm:
                if self.save_selection:
                    put(self.e, b, selection, False)
                if self.save_usage:
                    put(self.use, b, torch.zeros(B, ne), False)
                    put(self.life, b, torch.zeros(B, ne) + 1e-7, False)

    def update_usage(self, b, usage):
        if not self.save_usage:
            return
        usage = usage[:, self.perm_end.get(b, 0):]
        if usage.shape[-1] == 0:
            return
        self.use[b] = self.use[b] + usage
        self.life[b] = self.life[b] + 1

    def drop_oldest_keep(self, b, keep: int, min_size: int):
        """sieve_by_range(b, 0, -keep, min_size)."""
        p = self.perm_end.get(b, 0)
        n = self.size(b) - p
        if n <= min_size:
            return
        cut = n - keep if keep > 0 else n

        def sl(t, off):
            return torch.cat([t[..., :off], t[..., off + cut:]], -1)
        self.k[b], self.s[b] = sl(self.k[b], p), sl(self.s[b], p)
        if self.save_selection:
            self.e[b] = sl(self.e[b], 0)
        if self.save_usage:
            self.use[b], self.life[b] = sl(self.use[b], 0), sl(self.life[b], 0)
        for o in self.buckets[b]:
            self.v[o] = sl(self.v[o], p)

    def remove_obsolete(self, b, max_size):
        usage = self.use[b] / self.life[b]
        keep = [torch.topk(usage[bi], k=max_size)[1] for bi in range(usage.shape[0])]

        def g(t):
            return torch.stack([t[bi][..., keep[bi]] for bi in range(len(keep))], 0)
        self.k[b], self.s[b] = g(self.k[b]), g(self.s[b])
        for o in self.buckets[b]:
            self.v[o] = g(self.v[o])
        self.use[b], self.life[b] = g(self.use[b]), g(self.life[b])

    def purge_except(self, keep):
        for b in list(self.buckets):
            self.buckets[b] = [o for o in self.buckets[b] if o in keep]
            if not self.buckets[b]:
                for d in (self.buckets, self.k, self.s, self.e, self.use, self.life):
                    d.pop(b, None)
        self.v = {o: t for o, t in self.v.items() if o in keep}


class OracleCore:
    def __init__(self, network, cfg):
        self.net, self.cfg = network, cfg
        self.sd = {k: v.detach().float().cpu() for k, v in network.state_dict().items()}
        self.mem_every = cfg.mem_every
        self.flip_aug = cfg.flip_aug
        self.max_internal_size = cfg.max_internal_size
        self.top_k = cfg.top_k
        self.use_long_term = cfg.use_long_term
        st = cfg.stagger_updates
        self.stagger_ti = set(range(1, self.mem_every + 1)) if st >= self.mem_every else \
            set(np.round(np.linspace(1, self.mem_every, st)).astype(int))
        if self.use_long_term:
            lt = cfg.long_term
            self.max_mem_frames, self.min_mem_frames = lt.max_mem_frames - 1, lt.min_mem_frames - 1
            self.num_prototypes, self.max_long, self.buffer = lt.num_prototypes, lt.max_num_tokens, lt.buffer_tokens
            self.count_long_usage = lt.count_usage
        else:
            self.max_mem_frames = cfg.max_mem_frames - 1
        self.work = _Store(self.use_long_term, self.use_long_term)
        self.long = _Store(False, self.use_long_term and cfg.long_term.count_usage)
        self.sensory: Dict[int, torch.Tensor] = {}
        self.obj_v: Dict[int, torch.Tensor] = {}
        self.objects: List[int] = []           # live object ids in tmp-id order
        self.curr_ti, self.last_mem_ti = -1, 0
        self.last_mask = None
        self.last_logits = None
        self.engaged = False
        self.HW = None
        self.read_trace = None                 # filled with (idx, weights, usage) of the last read when set to {}
        self.selection_hook = None
        self.fg_hook = None                    # test-only, see oracle/transformer.query_transformer
        self.chunk_size = cfg.chunk_size

    # -- memory read (memory_manager.py:112-208) -----------------------------------------------
    def _read(self, pix_feat, key, selection):
        B, _, h, w = pix_feat.shape
        qk, qe = key.flatten(2), selection.flatten(2)
        out = {}
        for b, objs in self.work.buckets.items():
            long_n = self.long.size(b) if (self.use_long_term and b in self.long.buckets) else 0
            if long_n:
                mk = torch.cat([self.long.k[b], self.work.k[b]], -1)
                ms = torch.cat([self.long.s[b], self.work.s[b]], -1)
            else:
                mk, ms = self.work.k[b], self.work.s[b]
            sim = mm.similarity_expanded(mk, ms, qk, qe)
            idx, wts = mm.topk_softmax(sim, self.top_k)
            if self.selection_hook is not None:
                # test-only: lets a checker substitute an equally valid selection on near-tie queries
                # (SURVEY.md Appendix B take-away 5); weights are still exp(S)/sum exp(S) of THIS similarity
                idx = self.selection_hook(b, mk, ms, qk, qe, sim, idx)
                e = torch.gather(sim, 1, idx).exp()
                wts = e / e.sum(dim=1, keepdim=True)
            aff = mm.scatter_affinity(idx, wts, mk.shape[-1])
            if self.read_trace is not None:
                self.read_trace[b] = dict(idx=idx, weights=wts, sim=sim, affinity=aff)
            if self.use_long_term:
                usage = mm.usage_from_affinity(aff)
                self.work.update_usage(b, usage[:, long_n:])
                if long_n and self.count_long_usage:
                    self.long.update_usage(b, usage[:, :long_n])
            cs = self.chunk_size
            chunks = [objs] if cs < 1 else [objs[i:i + cs] for i in range(0, len(objs), cs)]
            for chunk in chunks:                                  # memory_manager.py:176-206
                vals = torch.stack([self.work.v[o] for o in chunk], 1)
                if long_n:
                    vals = torch.cat([torch.stack([self.long.v[o] for o in chunk], 1), vals], -1)
                visual = mm.readout(aff, vals).view(B, len(chunk), -1, h, w)
                sens = torch.stack([self.sensory[o] for o in chunk], 1)
                lm = self.last_mask[:, [self.objects.index(o) for o in chunk]]
                fused = self.net.pixel_fusion(pix_feat, visual, sens, lm)
                obj_mem = torch.stack([self.obj_v[o] for o in chunk], 1).unsqueeze(2)
                pix, _ = query_transformer(fused, obj_mem, self.sd, fg_hook=self.fg_hook)
                for i, o in enumerate(chunk):
                    out[o] = pix[:, i]
        return out

    # -- memory write (memory_manager.py:210-296) ----------------------------------------------
    def _add_memory(self, image, pix_feat, prob, key, shrinkage, selection, force_permanent):
        if prob.shape[1] == 0:
            return
        for o in self.objects:
            if o not in self.sensory:
                self.sensory[o] = torch.zeros(key.shape[0], self.cfg.model.sensory_dim, *key.shape[-2:])
        sens = torch.stack([self.sensory[o] for o in self.objects], 1)
        value, new_sens, summaries, _ = self.net.encode_mask(image, pix_feat, sens, prob, chunk_size=self.chunk_size)
        self.engaged = True
        if self.HW is None:
            self.HW = value.shape[-1] * value.shape[-2]
        k, s = key.flatten(2), shrinkage.flatten(2)
        e = selection.flatten(2) if selection is not None else None
        vflat = value.flatten(3)
        for i, o in enumerate(self.objects):
            if o in self.obj_v:
                self.obj_v[o] = self.obj_v[o] + summaries[:, i]
            else:
                self.obj_v[o] = summaries[:, i].clone()
        self.work.add(k, {o: vflat[:, i] for i, o in enumerate(self.objects)}, s, e,
                      as_permanent='all' if force_permanent else 'first')
        max_work = self.max_mem_frames * self.HW
        for b in list(self.work.buckets):
            if self.use_long_term:
                if self.work.non_perm(b) >= max_work:
                    if b in self.long.buckets and self.long.non_perm(b) >= self.max_long - self.num_prototypes:
                        self.long.remove_obsolete(b, self.max_long - self.num_prototypes - self.buffer)
                    self._compress(b)
            else:
                self.work.drop_oldest_keep(b, max_work, max_work)
        self.last_mem_ti = self.curr_ti
        for i, o in enumerate(self.objects):
            self.sensory[o] = new_sens[:, i]

    def _compress(self, b):
        """memory_manager.py:309-327."""
        min_work = self.min_mem_frames * self.HW
        p = self.work.perm_end.get(b, 0)
        end = -min_work if min_work > 0 else None
        ck, cs = self.work.k[b][:, :, p:end], self.work.s[b][:, :, p:end]
        ce = self.work.e[b][:, :, :end]
        cv = {o: self.work.v[o][:, :, p:end] for o in self.work.buckets[b]}
        usage = (self.work.use[b] / self.work.life[b])[:, :end]
        pk, pv, ps, _ = mm.consolidate(ck, cs, ce, cv, usage, self.num_prototypes)
        self.work.drop_oldest_keep(b, min_work, min_work)
        self.long.add(pk, pv, ps, None, supposed_bucket=b)

    # -- one frame (inference_core.py:172-328) --------------------------------------------------
    def step(self, image, mask=None, objects: Optional[List[int]] = None, *, idx_mask=True, end=False,
             force_permanent=False):
        if objects is None and mask is not None:
            objects = list(range(1, mask.shape[0] + 1))
        resize = False
        if self.max_internal_size > 0:
            h0, w0 = image.shape[-2:]
            short = min(h0, w0)
            if short > self.max_internal_size:
                resize = True
                nh, nw = int(h0 / short * self.max_internal_size), int(w0 / short * self.max_internal_size)
                image = F.interpolate(image[None], size=(nh, nw), mode='bilinear', align_corners=False)[0]
                if mask is not None:
                    if idx_mask:
                        mask = F.interpolate(mask[None, None].float(), size=(nh, nw),
                                             mode='nearest-exact')[0, 0].round().long()
                    else:
                        mask = F.interpolate(mask[None], size=(nh, nw), mode='bilinear', align_corners=False)[0]
        self.curr_ti += 1
        image, pad = _pad16(image)
        image = image[None]
        if self.flip_aug:
            image = torch.cat([image, image.flip(-1)], 0)
        since = self.curr_ti - self.last_mem_ti
        is_mem = (since >= self.mem_every or mask is not None) and not end
        need_seg = mask is None or (len(self.objects) > 0 and not all(o in self.objects for o in objects))
        upd_sens = (since in self.stagger_ti) and not end

        ms, pix_feat = self.net.encode_image(image)
        key, shrinkage, selection = self.net.transform_key(ms[0])
        if need_seg:
            if not self.engaged:
                prob = torch.zeros(1, key.shape[-2] * 16, key.shape[-1] * 16)
            else:
                ro = self._read(pix_feat, key, selection)
                ro = torch.stack([ro[o] for o in self.objects], 1)
                sens = torch.stack([self.sensory[o] for o in self.objects], 1)
                new_sens, logits, prob = self.net.segment(ms, ro, sens, chunk_size=self.chunk_size,
                                                             update_sensory=upd_sens)
                self.last_logits = logits
                prob = (prob[0] + prob[1].flip(-1)) / 2 if self.flip_aug else prob[0]
                if upd_sens:
                    for i, o in enumerate(self.objects):
                        self.sensory[o] = new_sens[:, i]
        if mask is not None:
            tmp_ids = []
            for o in objects:
                if o not in self.objects:
                    self.objects.append(o)
                tmp_ids.append(self.objects.index(o) + 1)
            mask, _ = _pad16(mask)
            if need_seg:
                nobg = prob[1:]
                if idx_mask:
                    nobg[:, mask > 0] = 0
                else:
                    nobg[:, mask.max(0) > 0.5] = 0
                extra = []
                for pos, t in enumerate(tmp_ids):
                    plane = (mask == objects[pos]).type_as(nobg) if idx_mask else mask[t]
                    if t > nobg.shape[0]:
                        extra.append(plane[None])
                    else:
                        nobg[t - 1] = plane
                mask = torch.cat([nobg, *extra], 0)
            elif idx_mask:
                if len(objects) == 0:
                    return torch.zeros(1, key.shape[-2] * 16, key.shape[-1] * 16)
                mask = torch.stack([mask == objects[i] for i in range(len(tmp_ids))], 0)
            prob = torch.softmax(aggregate_logits(mask.float(), dim=0), dim=0)
        self.last_mask = prob[1:][None]
        if self.flip_aug:
            self.last_mask = torch.cat([self.last_mask, self.last_mask.flip(-1)], 0)
        if is_mem or force_permanent:
            self._add_memory(image, pix_feat, self.last_mask, key, shrinkage, selection, force_permanent)
        lw, uw, lh, uh = pad
        H, W = prob.shape[-2:]
        out = prob[..., lh:H - uh, lw:W - uw]
        if resize:
            out = F.interpolate(out[None], size=(h0, w0), mode='bilinear', align_corners=False)[0]
        return out

    def delete_objects(self, objs: List[int]):
        """inference_core.py:330-335 + memory_manager.py:298-307 (obj_v deliberately kept)."""
        self.objects = [o for o in self.objects if o not in objs]
        self.work.purge_except(self.objects)
        self.long.purge_except(self.objects)
        self.sensory = {o: t for o, t in self.sensory.items() if o in self.objects}
        if not self.work.buckets:
            self.engaged = False

    def output_prob_to_mask(self, prob):
        m = prob.argmax(0)
        out = torch.zeros_like(m)
        for i, o in enumerate(self.objects):
            out[m == i + 1] = o
        return out
