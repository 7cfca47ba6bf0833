"""Runs the UNMODIFIED reference (baseline/_ref/ on the GPU box, /root/reference in the build container) on a clip in a
CHILD process -- the reference's package is also called `cutie`, so it cannot share an interpreter with this repo's
drop-in shim -- on a chosen device in eager fp32 with TF32 off (`amp=False`, eval_config.yaml:13), and hands back

  * the `CUTIE.segment(...)[1]` logits of every propagated frame (the parity tensor of SURVEY.md section 8 row a2),
  * the output masks,
  * the reference's complete recurrent state BEFORE every frame (working / long-term memory, usage counters, sensory
    memory, object summaries, last mask, frame clocks) in the attribute layout of oracle.cpu_core.OracleCore, so that
    tests/state_sync.load_state_from_oracle can teacher-force the product with it.

TEST INFRASTRUCTURE ONLY.  Parent side: `run_reference_clip(...)`; child side: `python tests/ref_runner.py in.pt out.pt`.
"""
import os
import subprocess
import sys
import tempfile
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    for cand in (os.environ.get('CUTIE_REFERENCE_ROOT'), '/root/reference', os.path.join(ROOT, 'baseline', '_ref')):
        if cand and os.path.isdir(os.path.join(cand, 'cutie', 'inference')):
            return cand
    return None


def run_reference_clip(frames, mask, objects, *, device='cuda', cfg_overrides=None, max_internal_size=-1,
                       snapshot=True, exact_similarity=False, timeout=900):
    """frames: list of [3,H,W] float CPU tensors; mask: index mask [H,W]; returns the child's result dict.

    exact_similarity=True is an ATTRIBUTION aid, not the reference as shipped: the one function
    `memory_utils.get_similarity` (bound into memory_manager, :9) is replaced by the same quantity evaluated in float64
    direct form, -ms/sqrt(CK) * sum_c qe (mk-qk)^2, rounded to fp32.  The reference's own fp32 three-term expansion
    (-a^2 + 2ab - b^2, memory_utils.py:28-36) cancels catastrophically, so its top-k selection on near-tied queries is
    decided by GEMM rounding noise (cuBLAS vs MKL pick differently: SURVEY.md section 7 "top-k fidelity"); with the noise
    removed, whatever still differs from the CUDA path is not selection noise."""
    root = reference_root()
    if root is None:
        raise RuntimeError('no reference tree (baseline/_ref missing: run `python baseline/install_reference.py`)')
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, 'in.pt'), os.path.join(tmp, 'out.pt')
        torch.save(dict(frames=[f.cpu() for f in frames], mask=mask.cpu(), objects=list(objects), device=device,
                        cfg_overrides=dict(cfg_overrides or {}), max_internal_size=max_internal_size,
                        snapshot=snapshot, exact_similarity=bool(exact_similarity)), fin)
        env = dict(os.environ, CUTIE_REFERENCE_ROOT=root)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), fin, fout], env=env, cwd=ROOT,
                           capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError('reference child failed:\n' + r.stdout[-2000:] + '\n' + r.stderr[-4000:])
        return torch.load(fout, weights_only=False)


def as_oracle_state(snap):
    """dict snapshot -> an object with OracleCore's state attributes (what load_state_from_oracle reads)."""
    def store(d):
        return types.SimpleNamespace(**d)
    return types.SimpleNamespace(objects=snap['objects'], curr_ti=snap['curr_ti'], last_mem_ti=snap['last_mem_ti'],
                                 last_mask=snap['last_mask'], work=store(snap['work']), long=store(snap['long']),
                                 sensory=snap['sensory'], obj_v=snap['obj_v'], engaged=snap['engaged'])


# ----------------------------------------------------------------------------------------------------------
# child
def _cpu(t):
    return t.detach().float().cpu().clone()


def _snap_store(st):
    if st is None:
        return dict(buckets={}, perm_end={}, k={}, s={}, v={}, e={}, use={}, life={}, next_bucket=0)
    d = dict(buckets={b: [int(o) for o in objs] for b, objs in st.buckets.items()},
             perm_end={b: int(p) for b, p in st.perm_end_pt.items()},
             k={b: _cpu(t) for b, t in st.k.items()}, s={b: _cpu(t) for b, t in st.s.items()},
             v={int(o): _cpu(t) for o, t in st.v.items()}, e={}, use={}, life={},
             next_bucket=int(st.global_bucket_id))
    if st.save_selection:
        d['e'] = {b: _cpu(t) for b, t in st.e.items()}
    if st.save_usage:
        d['use'] = {b: _cpu(t) for b, t in st.use_cnt.items()}
        d['life'] = {b: _cpu(t) for b, t in st.life_cnt.items()}
    return d


def _snapshot(proc):
    m = proc.memory
    return dict(objects=[int(o) for o in proc.object_manager.all_obj_ids], curr_ti=proc.curr_ti,
                last_mem_ti=proc.last_mem_ti, last_mask=None if proc.last_mask is None else _cpu(proc.last_mask),
                work=_snap_store(m.work_mem), long=_snap_store(m.long_mem if m.use_long_term else None),
                sensory={int(o): _cpu(t) for o, t in m.sensory.items()},
                obj_v={int(o): _cpu(t) for o, t in m.obj_v.items()}, engaged=bool(m.engaged))


def _child(fin, fout):
    sys.path.insert(0, ROOT)
    import warnings
    warnings.filterwarnings('ignore')
    from oracle import ref_harness as rh
    job = torch.load(fin, weights_only=False)
    dev = torch.device(job['device'])
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    ref = rh.load_reference()
    cfg = rh.reference_cfg(**job['cfg_overrides'])
    net = rh.build_reference_model(cfg).to(dev)
    if job.get('exact_similarity'):
        import math

        def similarity_f64(mk, ms, qk, qe, add_batch_dim=False):
            if add_batch_dim:
                mk, ms, qk, qe = mk.unsqueeze(0), ms.unsqueeze(0), qk.unsqueeze(0), qe.unsqueeze(0)
            CK = mk.shape[1]
            mk, qk = mk.flatten(2).double(), qk.flatten(2).double()
            qe = qe.flatten(2).double() if qe is not None else torch.ones_like(qk)
            out = torch.empty(mk.shape[0], mk.shape[2], qk.shape[2], dtype=torch.float32, device=mk.device)
            for n0 in range(0, mk.shape[2], 256):
                d = mk[:, :, n0:n0 + 256, None] - qk[:, :, None, :]
                out[:, n0:n0 + 256] = (-(d * d * qe[:, :, None, :]).sum(1) / math.sqrt(CK)).float()
            if ms is not None:
                out = (out.double() * ms.flatten(1).unsqueeze(2).double()).float()
            return out
        ref.memory_manager.get_similarity = similarity_f64
    proc = ref.InferenceCore(net, cfg=cfg)
    if job['max_internal_size'] > 0:
        proc.max_internal_size = job['max_internal_size']
    logits, masks, states = [], [], []
    orig = net.segment

    def seg(*a, **kw):
        s, lg, p = orig(*a, **kw)
        logits.append(_cpu(lg))
        return s, lg, p
    net.segment = seg
    mask = job['mask'].to(dev)
    with torch.inference_mode():
        for ti, f in enumerate(job['frames']):
            if job['snapshot']:
                states.append(_snapshot(proc))
            n0 = len(logits)
            prob = proc.step(f.to(dev), mask, objects=job['objects']) if ti == 0 else proc.step(f.to(dev))
            if len(logits) == n0:
                logits.append(None)
            masks.append(proc.output_prob_to_mask(prob).cpu())
    torch.save(dict(logits=logits, masks=masks, states=states, root=rh.REF_ROOT, device=str(dev),
                    torch=torch.__version__), fout)


if __name__ == '__main__':
    _child(sys.argv[1], sys.argv[2])
