import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cutie_b200.config import default_config
from cutie_b200.model.cutie import CUTIE
from oracle.synth import synthetic_state_dict
g = np.load('tests/golden/qt_module.npz')
cfg = default_config()
net = CUTIE(cfg).eval(); net.load_state_dict(synthetic_state_dict(net.state_dict(), 0))
qt = net.object_transformer.cuda()
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    with torch.inference_mode():
        out, aux = qt(torch.from_numpy(g['pixel']).cuda(), torch.from_numpy(g['obj_summaries']).cuda())
    ref = torch.from_numpy(g['out'])
    d = (out.cpu() - ref).abs()
    print('cudnn tf32', tf32, 'max diff', float(d.max()), 'frac>1e-3', float((d > 1e-3).float().mean()),
          'aux0', float((aux['logits'][0].cpu() - torch.from_numpy(g['aux_logits_0'])).abs().max()),
          'aux3', float((aux['logits'][3].cpu() - torch.from_numpy(g['aux_logits_3'])).abs().max()))
