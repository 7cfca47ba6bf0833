"""Diagnostic (GPU box): per-frame logit differences, free-running GPU path vs CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from cutie_b200.config import default_config
from cutie_b200.inference.inference_core import InferenceCore
from cutie_b200.model.cutie import CUTIE
from oracle.cpu_core import OracleCore
from oracle.synth import synthetic_state_dict, synthetic_video

def net(cfg, cuda):
    n = CUTIE(cfg).eval(); n.load_state_dict(synthetic_state_dict(n.state_dict(), 0))
    return n.cuda() if cuda else n

H, W, K, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
cfg = default_config(mem_every=3, max_mem_frames=4)
proc, oc = InferenceCore(net(cfg, True), cfg=cfg), OracleCore(net(cfg, False), cfg)
frames, mask = synthetic_video(T, H, W, K, seed=5)
objs = list(range(1, K + 1))
with torch.inference_mode():
    for ti in range(T):
        if ti == 0:
            proc.step(frames[0].cuda(), mask.cuda(), objects=objs); oc.step(frames[0], mask, objects=objs)
        else:
            proc.step(frames[ti].cuda()); oc.step(frames[ti])
            d = (proc.last_logits.cpu() - oc.last_logits).abs()
            sd = max(float((proc.memory.sensory[o].cpu() - oc.sensory[o]).abs().max()) for o in objs)
            print(f'frame {ti}: max {float(d.max()):.3e}  frac>1e-3 {float((d > 1e-3).float().mean()):.2e}  sensory diff {sd:.2e}', flush=True)
