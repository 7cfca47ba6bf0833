"""BASELINE.json configs[0]: the reference's own scripting_demo.py case (examples/images/bike, 854x480, the mask file's
two labels) run through the UNMODIFIED reference on CPU, fp32, following scripting_demo.py:17-58 (get_default_model's
config, max_internal_size = 480, first-frame mask memorised, three propagated frames).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_bike.py
Writes tests/golden/cfg1_bike.npz: the four input JPEGs and the mask PNG as raw file bytes (the GPU box has no
/root/reference; PIL decodes them there exactly as here), the reference's segment() logits of the propagated frames at
every 4th pixel (offset 2: the decoder's native stride), its output masks and mean probabilities.  Weights are the
name-seeded synthetic ones (cutie_b200/utils/synth.py) -- real checkpoints are unobtainable offline."""
import io
import os
import sys
import warnings

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')

from oracle import ref_harness as rh          # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
IMG_DIR = os.path.join(rh.REF_ROOT, 'examples', 'images', 'bike')
MASK = os.path.join(rh.REF_ROOT, 'examples', 'masks', 'bike', '00000.png')


def load_inputs(jpegs, mask_png):
    """Decodes exactly as scripting_demo.py does (PIL -> to_tensor == uint8 HWC / 255 -> CHW float)."""
    frames = [torch.from_numpy(np.array(Image.open(io.BytesIO(bytes(b))).convert('RGB'))).permute(2, 0, 1).float() / 255
              for b in jpegs]
    m = Image.open(io.BytesIO(bytes(mask_png)))
    assert m.mode in ('L', 'P')
    mask = torch.from_numpy(np.array(m))
    objects = [int(o) for o in np.unique(np.array(m)) if o != 0]
    return frames, mask, objects


def main():
    names = sorted(os.listdir(IMG_DIR))
    jpegs = [np.frombuffer(open(os.path.join(IMG_DIR, n), 'rb').read(), dtype=np.uint8) for n in names]
    mask_png = np.frombuffer(open(MASK, 'rb').read(), dtype=np.uint8)
    frames, mask, objects = load_inputs(jpegs, mask_png)
    ref = rh.load_reference()
    cfg = rh.reference_cfg()                       # eval_config.yaml + model/base.yaml, mem_every=5 (get_default_model)
    net = rh.build_reference_model(cfg)
    proc = ref.InferenceCore(net, cfg=cfg)
    proc.max_internal_size = 480                   # scripting_demo.py:22
    logits, masks, probs = [], [], []
    orig = net.segment

    def seg(*a, **kw):
        s, lg, p = orig(*a, **kw)
        logits.append(lg.detach().clone())
        return s, lg, p
    net.segment = seg
    with torch.inference_mode():
        for ti, f in enumerate(frames):
            prob = proc.step(f, mask, objects=objects) if ti == 0 else proc.step(f)
            masks.append(proc.output_prob_to_mask(prob).numpy().astype(np.uint8))
            probs.append(prob.mean(dim=(1, 2)).numpy())
    net.segment = orig
    lg = torch.cat(logits, 0)                      # [3, 1+K, Hp, Wp]
    print('objects', objects, 'logits', tuple(lg.shape), 'range', float(lg.min()), float(lg.max()))
    np.savez_compressed(os.path.join(OUT, 'cfg1_bike.npz'),
                        **{f'jpeg_{i}': j for i, j in enumerate(jpegs)}, mask_png=mask_png,
                        objects=np.array(objects), logits_s4=lg[:, :, 2::4, 2::4].numpy().astype(np.float32),
                        logits_shape=np.array(lg.shape), masks=np.stack(masks, 0), mean_prob=np.stack(probs, 0))
    print('wrote cfg1_bike.npz')


if __name__ == '__main__':
    main()
