"""Encoders / decoder / fuser that sit either side of the hot path -- PyTorch + cuDNN, unchanged in role.

Fresh restatements with the reference's state-dict names (cutie/model/big_modules.py:21-306).
"""
from typing import List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from cutie_b200 import kernels as K_
from cutie_b200.model.backbone import ResNetTrunk
from cutie_b200.model.fuse import conv_relu_maxpool
from cutie_b200.model.blocks import (DeepSensoryUpdater, FeatureFusion, MultiScaleSensoryUpdater, ObjConv2d,
                                     UpsampleBlock, fold, unfold)


def _chunks(n: int, size: int):
    if size < 1 or size >= n:
        return [(0, n)]
    return [(i, min(i + size, n)) for i in range(0, n, size)]


class PixelEncoder(nn.Module):
    """big_modules.py:21-61: image -> (f16, f8, f4)."""

    def __init__(self, model_cfg):
        super().__init__()
        trunk = ResNetTrunk(model_cfg.pixel_encoder.type)
        self.conv1, self.bn1 = trunk.conv1, trunk.bn1
        self.res2, self.layer2, self.layer3 = trunk.layer1, trunk.layer2, trunk.layer3
        self.channels_last = False

    def forward(self, x):
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        if getattr(self, 'bn_folded', False):
            x = conv_relu_maxpool(self.conv1, x)
        else:
            x = F.max_pool2d(F.relu(self.bn1(self.conv1(x)), inplace=True), 3, stride=2, padding=1)
        f4 = self.res2(x)
        f8 = self.layer2(f4)
        return self.layer3(f8), f8, f4

    def train(self, mode: bool = True):
        return super().train(False)     # BN statistics are frozen (big_modules.py:55-61)


class KeyProjection(nn.Module):
    """big_modules.py:64-87: f16 -> key [64], shrinkage d^2+1 [1], selection sigmoid [64]."""

    def __init__(self, model_cfg):
        super().__init__()
        c_in, mid, ck = model_cfg.pixel_encoder.ms_dims[0], model_cfg.pixel_dim, model_cfg.key_dim
        self.pix_feat_proj = nn.Conv2d(c_in, mid, 1)
        self.key_proj = nn.Conv2d(mid, ck, 3, padding=1)
        self.d_proj = nn.Conv2d(mid, 1, 3, padding=1)
        self.e_proj = nn.Conv2d(mid, ck, 3, padding=1)

    def forward(self, f16, *, need_s: bool, need_e: bool):
        x = self.pix_feat_proj(f16)
        shrinkage = None
        if need_s:
            t = getattr(self, 'glue_dispatch', None)
            if t is None:
                d = self.d_proj(x)
            else:   # 3x3 with ONE output channel: cutie_conv3x3_c1 (cuDNN: a 143 us GEMM with N = 1 at 480p)
                d = t('pred_conv3x3', (tuple(x.shape),), lambda: self.d_proj(x),
                      lambda trial: K_.conv3x3_c1(x.contiguous(), self.d_proj.weight, self.d_proj.bias), x)
            shrinkage = d.square() + 1
        selection = torch.sigmoid(self.e_proj(x)) if need_e else None
        return self.key_proj(x), shrinkage, selection


class MaskEncoder(nn.Module):
    """big_modules.py:90-189: (image, pix_feat, sensory, masks, others) -> (value [B,K,CV,h,w], new sensory)."""

    def __init__(self, model_cfg, single_object: bool = False):
        super().__init__()
        self.single_object = single_object
        trunk = ResNetTrunk(model_cfg.mask_encoder.type, extra_in=1 if single_object else 2)
        self.conv1, self.bn1 = trunk.conv1, trunk.bn1
        self.layer1, self.layer2, self.layer3 = trunk.layer1, trunk.layer2, trunk.layer3
        self.fuser = FeatureFusion(model_cfg.pixel_dim, model_cfg.mask_encoder.final_dim, model_cfg.value_dim)
        self.sensory_update = DeepSensoryUpdater(model_cfg.value_dim, model_cfg.sensory_dim)
        self.channels_last = False

    def forward(self, image, pix_feat, sensory, masks, others, *, deep_update=True, chunk_size=-1):
        B, K = masks.shape[:2]
        extra = masks.unsqueeze(2) if self.single_object else torch.stack([masks, others], 2)
        g = torch.cat([image.unsqueeze(1).expand(-1, K, -1, -1, -1), extra], 2)    # [B,K,3+e,H,W]
        spans = _chunks(K, chunk_size)
        new_sensory = sensory if (len(spans) == 1 or not deep_update) else torch.empty_like(sensory)
        values = []
        for lo, hi in spans:
            t = fold(g[:, lo:hi])
            if self.channels_last:
                t = t.contiguous(memory_format=torch.channels_last)
            if getattr(self, 'bn_folded', False):
                # relu and max-pool commute (both monotone): relu(maxpool(y)) == maxpool(relu(y)), bit for bit
                t = conv_relu_maxpool(self.conv1, t)
            else:
                t = F.relu(F.max_pool2d(self.bn1(self.conv1(t)), 3, stride=2, padding=1))
            t = self.layer3(self.layer2(self.layer1(t)))
            if self.channels_last:
                t = t.contiguous()
            v = self.fuser(pix_feat, unfold(t, B))
            values.append(v)
            if deep_update:
                upd = self.sensory_update(v, sensory[:, lo:hi])
                if len(spans) == 1:
                    new_sensory = upd
                else:
                    new_sensory[:, lo:hi] = upd
        return torch.cat(values, 1), new_sensory

    def train(self, mode: bool = True):
        return super().train(False)


class PixelFeatureFuser(nn.Module):
    """big_modules.py:192-235: (pix_feat, memory readout, sensory, last mask) -> fused [B,K,E,h,w]."""

    def __init__(self, model_cfg, single_object: bool = False):
        super().__init__()
        self.single_object = single_object
        self.fuser = FeatureFusion(model_cfg.pixel_dim, model_cfg.value_dim, model_cfg.embed_dim)
        self.sensory_compress = ObjConv2d(model_cfg.sensory_dim + (1 if single_object else 2),
                                          model_cfg.value_dim, 1)

    def forward(self, pix_feat, pixel_memory, sensory_memory, last_mask, last_others, *, chunk_size=-1):
        K = pixel_memory.shape[1]
        m = last_mask.unsqueeze(2) if self.single_object else torch.stack([last_mask, last_others], 2)
        outs = []
        for lo, hi in _chunks(K, chunk_size):
            s = self.sensory_compress(torch.cat([sensory_memory[:, lo:hi], m[:, lo:hi]], 2))
            outs.append(self.fuser(pix_feat, pixel_memory[:, lo:hi] + s))
        return outs[0] if len(outs) == 1 else torch.cat(outs, 1)


class _SkipProjections(nn.Module):
    def __init__(self, in_dims: List[int], out_dims: List[int]):
        super().__init__()
        self.transforms = nn.ModuleList([nn.Conv2d(i, o, 1) for i, o in zip(in_dims, out_dims)])

    def forward(self, feats):
        return [t(f) for t, f in zip(self.transforms, feats)]


class MaskDecoder(nn.Module):
    """big_modules.py:238-306: (ms features, memory readout, sensory) -> (new sensory, logits [B,K,4h,4w])."""

    def __init__(self, model_cfg):
        super().__init__()
        up = model_cfg.mask_decoder.up_dims
        sd = model_cfg.sensory_dim
        assert model_cfg.embed_dim == up[0]
        self.sensory_update = MultiScaleSensoryUpdater([up[0], up[1], up[2] + 1], sd, sd)
        self.decoder_feat_proc = _SkipProjections(model_cfg.pixel_encoder.ms_dims[1:], up[:-1])
        self.up_16_8 = UpsampleBlock(up[0], up[1])
        self.up_8_4 = UpsampleBlock(up[1], up[2])
        self.pred = nn.Conv2d(up[-1], 1, 3, padding=1)

    def forward(self, ms_image_feat, memory_readout, sensory, *, chunk_size=-1, update_sensory=True):
        B, K = memory_readout.shape[:2]
        f8, f4 = self.decoder_feat_proc(ms_image_feat[1:])
        spans = _chunks(K, chunk_size)
        new_sensory = sensory if (len(spans) == 1 or not update_sensory) else torch.empty_like(sensory)
        all_logits = []
        for lo, hi in spans:
            p16 = memory_readout[:, lo:hi]
            p8 = self.up_16_8(p16, f8)
            p4 = self.up_8_4(p8, f4)
            x4 = fold(p4).float()

            def aten():
                return self.pred(F.relu(x4))
            t = getattr(self, 'glue_dispatch', None)
            if t is None:
                lg = aten()
            else:   # single-output-channel 3x3 convolution: cutie_conv3x3_c1 (ReLU on the fly, no cuDNN transposes)
                tol = 2e-2 if torch.backends.cudnn.allow_tf32 else 2e-4
                lg = t('pred_conv3x3', (tuple(x4.shape),), aten,
                       lambda trial: K_.conv3x3_c1(x4, self.pred.weight, self.pred.bias, relu_input=True), x4, rtol=tol)
            lg = unfold(lg, B)                                              # [B,k,1,4h,4w]
            if update_sensory:
                upd = self.sensory_update(p16, p8, (p4, lg), sensory[:, lo:hi])
                if len(spans) == 1:
                    new_sensory = upd
                else:
                    new_sensory[:, lo:hi] = upd
            all_logits.append(lg.squeeze(2))
        return new_sensory, (all_logits[0] if len(all_logits) == 1 else torch.cat(all_logits, 1))
