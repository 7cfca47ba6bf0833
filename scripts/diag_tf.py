"""Diagnostic (GPU box): teacher-forced per-frame comparison with stage-level diffs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from cutie_b200.config import default_config
from cutie_b200.inference.inference_core import InferenceCore
from cutie_b200.model.cutie import CUTIE
import cutie_b200.kernels as K_
from oracle.cpu_core import OracleCore
from oracle.synth import synthetic_state_dict, synthetic_video
from tests.state_sync import load_state_from_oracle

def net(cfg, cuda):
    n = CUTIE(cfg).eval(); n.load_state_dict(synthetic_state_dict(n.state_dict(), 0))
    return n.cuda() if cuda else n

CASES = {
 'lt': (dict(mem_every=2, use_long_term=True, long_term=dict(max_mem_frames=4, min_mem_frames=2, num_prototypes=64, max_num_tokens=600, buffer_tokens=100)), 240, 432, 3, 16),
 'flip': (dict(mem_every=3, max_mem_frames=4, flip_aug=True, chunk_size=1), 240, 432, 3, 8),
 'k50': (dict(mem_every=2, max_mem_frames=3, top_k=50), 480, 854, 3, 5),
 'k30_480': (dict(mem_every=2, max_mem_frames=3), 480, 854, 3, 5),
}
over, H, W, K, T = CASES[sys.argv[1]]
cfg = default_config(**over)
proc, oc = InferenceCore(net(cfg, True), cfg=cfg), OracleCore(net(cfg, False), cfg)
frames, mask = synthetic_video(T, H, W, K, seed=5)
objs = list(range(1, K + 1))
orig = K_.affinity_topk
cap = {}
def spy(segs, qk, qe, top_k, usage_acc=None, want_sim=False):
    idx, w, sim = orig(segs, qk, qe, top_k, usage_acc=usage_acc, want_sim=True)
    cap['idx'], cap['w'], cap['sim'] = idx.cpu(), w.cpu(), sim.cpu()
    return idx, w, sim
K_.affinity_topk = spy
import cutie_b200.inference.memory_manager as MM
with torch.inference_mode():
    for ti in range(T):
        load_state_from_oracle(proc, oc, 'cuda')
        oc.read_trace = {}
        if ti == 0:
            pg = proc.step(frames[0].cuda(), mask.cuda(), objects=objs); pc = oc.step(frames[0], mask, objects=objs)
            print('frame 0 prob diff', float((pg.cpu() - pc).abs().max()))
            continue
        pg = proc.step(frames[ti].cuda()); pc = oc.step(frames[ti])
        d = (proc.last_logits.cpu() - oc.last_logits).abs()
        tr = oc.read_trace[0]
        k = cfg.top_k
        gi = cap['idx'][:, :, :k].transpose(1, 2).long().sort(1)[0]
        oi = tr['idx'].sort(1)[0]
        same = (gi == oi).all(1).float().mean()
        wdiff = float((cap['w'][:, :, :k].transpose(1, 2).sort(1)[0] - tr['weights'].sort(1)[0]).abs().max())
        print(f'frame {ti}: logits max {float(d.max()):.3e} frac>1e-3 {float((d > 1e-3).float().mean()):.2e} | topk sets equal {float(same):.4f} wdiff {wdiff:.2e} '
              f'| work {proc.memory.work_mem.size(0)}/{oc.work.size(0)} long {proc.memory.long_mem.size(0) if cfg.use_long_term else 0}/{oc.long.size(0)}', flush=True)
