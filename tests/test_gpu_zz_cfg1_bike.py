"""BASELINE.json configs[0] -- the reference's scripting_demo.py case (examples/images/bike: four 854x480 frames, the two
labels of the mask file) -- against the fixture the UNMODIFIED reference produced for it on CPU
(tests/golden/make_golden_bike.py -> cfg1_bike.npz).  CPU: the oracle's full-frame restatement; GPU: InferenceCore on the
fused kernels, eager and with CUDA graphs, driven exactly like scripting_demo.py:17-58.

Asserted: the memorised first frame and the FIRST propagated frame (logits within the bar, masks equal).  Frames 3 and 4
are run and reported but not asserted: with random-init weights (no checkpoint offline) the recurrent net amplifies
rounding differences ~3 000x per 480p frame (oracle vs reference, both fp32 on the CPU: 1e-5 -> 4e-2 -> 1.6), so a
free-running comparison says nothing beyond the first propagated frame; later frames are covered teacher-forced
(tests/test_gpu_e2e.py)."""
import io
import os

import numpy as np
import pytest
import torch
from PIL import Image

from tests.conftest import GOLDEN


def _inputs(g):
    frames = [torch.from_numpy(np.array(Image.open(io.BytesIO(g[f'jpeg_{i}'].tobytes())).convert('RGB')))
              .permute(2, 0, 1).float() / 255 for i in range(4)]
    mask = torch.from_numpy(np.array(Image.open(io.BytesIO(g['mask_png'].tobytes()))))
    return frames, mask, [int(o) for o in g['objects']]


def _net():
    from cutie_b200.config import default_config
    from cutie_b200.model.cutie import CUTIE
    from cutie_b200.utils.synth import synthetic_state_dict
    cfg = default_config()                       # == get_default_model(): eval_config.yaml + base.yaml, mem_every=5
    net = CUTIE(cfg).eval()
    net.load_state_dict(synthetic_state_dict(net.state_dict(), 0))
    return cfg, net


def _compare(g, logits, masks, tol):
    ref = torch.from_numpy(g['logits_s4'])
    got = torch.cat(logits, 0)
    assert tuple(got.shape) == tuple(int(x) for x in g['logits_shape'])
    per_frame = (got[:, :, 2::4, 2::4].cpu() - ref).abs().flatten(1).max(1)[0]
    print('max |logit diff| per propagated frame (only the first is asserted):', [float(x) for x in per_frame])
    assert float(per_frame[0]) < tol, float(per_frame[0])
    ref_masks = g['masks']
    for ti in (0, 1):
        differ = float((masks[ti].cpu().numpy().astype(np.uint8) != ref_masks[ti]).mean())
        assert differ < 2e-4, (ti, differ)       # a handful of boundary pixels may flip within the logit tolerance
    return float(per_frame[0])


def test_oracle_matches_the_reference_on_the_bike_example():
    from oracle.cpu_core import OracleCore
    g = np.load(os.path.join(GOLDEN, 'cfg1_bike.npz'))
    frames, mask, objects = _inputs(g)
    assert frames[0].shape == (3, 480, 854) and objects == [1, 2]
    cfg, net = _net()
    oc = OracleCore(net, cfg)
    logits, masks = [], []
    with torch.inference_mode():
        for ti, f in enumerate(frames):
            prob = oc.step(f, mask, objects=objects) if ti == 0 else oc.step(f)
            if ti > 0:
                logits.append(oc.last_logits.clone())
            masks.append(prob.argmax(0))
    # object ids == tmp ids here (labels 1, 2 in order), so argmax is the output mask
    err = _compare(g, logits, masks, 2e-4)
    print('oracle vs reference, bike: max |logit diff| =', err)


@pytest.mark.gpu
@pytest.mark.timeout(900, method='thread')
@pytest.mark.parametrize('graphs', [False, True])
def test_cuda_path_matches_the_reference_on_the_bike_example(graphs):
    from cutie_b200.inference.inference_core import InferenceCore
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(GOLDEN, 'cfg1_bike.npz'))
    frames, mask, objects = _inputs(g)
    cfg, net = _net()
    proc = InferenceCore(net.cuda(), cfg=cfg, use_cuda_graphs=graphs)
    proc.max_internal_size = 480                 # scripting_demo.py:22
    logits, masks = [], []
    with torch.inference_mode():
        for ti, f in enumerate(frames):
            prob = proc.step(f.cuda(), mask.cuda(), objects=objects) if ti == 0 else proc.step(f.cuda())
            if ti > 0:
                logits.append(proc.last_logits.clone())
            masks.append(proc.output_prob_to_mask(prob))
    err = _compare(g, logits, masks, 1e-3)
    print(f'CUDA path (graphs={graphs}) vs reference, bike: max |logit diff| =', err)
