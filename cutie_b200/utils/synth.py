"""Deterministic synthetic weights and inputs shared by the oracle, the tests and bench.py.

Data generation only -- no Cutie arithmetic lives here (oracle/synth.py re-exports it for the oracle and the tests, so
that bench.py's GPU arm imports nothing from oracle/).  TEST/BENCH INFRASTRUCTURE.  Real Cutie checkpoints (cutie-base-mega.pth) are downloaded from
GitHub by the reference (cutie/utils/download_models.py:8-11) and are unobtainable offline, and a
140 MB state_dict cannot be committed as a fixture.  Instead every tensor of a state_dict is filled
from a numpy PCG64 stream seeded by crc32(key-name) so that *any* module tree that exposes the same
state_dict key names and shapes (the reference's CUTIE here, cutie_b200's CUTIE on the GPU box)
receives bit-identical weights without sharing construction order.
"""
import zlib
import numpy as np
import torch


def _rng(name: str, seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([zlib.crc32(name.encode()), seed]))


# per-tensor gains that keep a random-weight network in the numeric regime of a trained one
# (keys O(1), shrinkage in [1, ~4], similarities O(-10) so the un-shifted exp of
# memory_utils.py:60 does not underflow to 0/0).
_GAIN = {
    'key_proj.pix_feat_proj.weight': 0.2,
    'key_proj.d_proj.weight': 0.6,
    'key_proj.key_proj.weight': 0.7,
    'pix_feat_proj.weight': 0.2,
    'mask_encoder.fuser.block2.conv2.weight': 0.3,
    'pixel_fuser.fuser.block2.conv2.weight': 0.2,
    'object_summarizer.feature_pred.2.weight': 0.3,
    'object_transformer.pixel_init_proj.weight': 0.5,
    'object_transformer.pixel_emb_proj.weight': 0.3,
    'object_transformer.summary_to_query_init.weight': 0.3,
    'object_transformer.summary_to_query_emb.weight': 0.3,
    'mask_decoder.up_16_8.out_conv.conv2.weight': 0.3,
    'mask_decoder.up_8_4.out_conv.conv2.weight': 0.3,
    'mask_decoder.pred.weight': 0.3,
}


def synthetic_state_dict(template: dict, seed: int = 0) -> dict:
    """template: name -> tensor (only shape/dtype are used).  Returns name -> new tensor."""
    out = {}
    for name, t in template.items():
        shape = tuple(t.shape)
        g = _rng(name, seed)
        leaf = name.rsplit('.', 1)[-1]
        if leaf == 'inv_freq':
            out[name] = t.detach().clone()  # analytic buffer (positional_encoding.py:30-31)
            continue
        if leaf == 'num_batches_tracked':
            out[name] = torch.zeros(shape, dtype=t.dtype)
            continue
        n = g.standard_normal(shape).astype(np.float32) if len(shape) else np.float32(0)
        is_norm = ('bn' in name.split('.')[-2] or 'norm' in name.split('.')[-2]
                   or 'downsample.1' in name) if '.' in name else False
        if leaf == 'running_mean':
            v = 0.05 * n
        elif leaf == 'running_var':
            v = 1.0 + 0.1 * np.abs(n)
        elif is_norm and leaf == 'weight':
            v = 1.0 + 0.05 * n
            if name.endswith('bn3.weight') or (name.endswith('bn2.weight') and 'mask_encoder' in name):
                v = 0.5 * v  # damp the residual branch so 16 stacked blocks stay O(1)
        elif is_norm and leaf == 'bias':
            v = 0.05 * n
        elif leaf == 'bias' or leaf == 'in_proj_bias':
            v = 0.05 * n
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            if 'query_init' in name or 'query_emb' in name:
                v = n  # nn.Embedding default N(0,1)
            elif len(shape) == 4:
                v = n * np.sqrt(2.0 / fan_in)
            else:
                v = n * np.sqrt(1.0 / fan_in)
        else:
            v = 0.05 * n
        v = v * _GAIN.get(name, 1.0)
        out[name] = torch.from_numpy(np.asarray(v, dtype=np.float32).reshape(shape)).to(t.dtype)
    return out


def synthetic_video(num_frames: int, H: int, W: int, num_objects: int, seed: int = 0):
    """Temporally coherent clip + first-frame index mask.

    frame_t = clamp(base + 0.05*N(0,1), 0, 1) around a smooth-ish base; the mask holds
    `num_objects` axis-aligned rectangles with ids 1..K.  Returns (frames [T,3,H,W] f32, mask [H,W] i64).
    """
    g = np.random.Generator(np.random.PCG64([0xC071E, seed]))
    coarse = g.random((3, max(H // 16, 2), max(W // 16, 2))).astype(np.float32)
    base = torch.nn.functional.interpolate(torch.from_numpy(coarse)[None], size=(H, W), mode='bilinear',
                                           align_corners=False)[0].numpy()
    base = 0.7 * base + 0.3 * g.random((3, H, W)).astype(np.float32)
    frames = np.empty((num_frames, 3, H, W), np.float32)
    for t in range(num_frames):
        frames[t] = np.clip(base + 0.05 * g.standard_normal((3, H, W)).astype(np.float32), 0, 1)
    mask = np.zeros((H, W), np.int64)
    for k in range(num_objects):
        rh, rw = max(H // 4, 2), max(W // (num_objects + 2), 2)
        y0 = (H // 6) + (k * H) // (3 * max(num_objects, 1))
        x0 = (W // (num_objects + 1)) * k + W // (4 * (num_objects + 1))
        mask[y0:min(y0 + rh, H), x0:min(x0 + rw, W)] = k + 1
    return torch.from_numpy(frames), torch.from_numpy(mask)
