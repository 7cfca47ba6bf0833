"""Object summaries written on memory frames (cutie/model/transformer/object_summarizer.py:26-88).

Runs once per `mem_every` frames; SURVEY.md section 8(f).4 ranks it "next", so it stays PyTorch here.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from cutie_b200.model.blocks import area_resize
from cutie_b200.model.positional import SinusoidPE


class ObjectSummarizer(nn.Module):
    def __init__(self, model_cfg):
        super().__init__()
        c = model_cfg.object_summarizer
        E = c.embed_dim
        self.num_summaries = c.num_summaries
        self.add_pe = c.add_pe
        if self.add_pe:
            self.pos_enc = SinusoidPE(E, model_cfg.pixel_pe_scale, model_cfg.pixel_pe_temperature)
        self.input_proj = nn.Linear(model_cfg.value_dim, E)
        self.feature_pred = nn.Sequential(nn.Linear(E, E), nn.ReLU(inplace=True), nn.Linear(E, E))
        self.weights_pred = nn.Sequential(nn.Linear(E, E), nn.ReLU(inplace=True), nn.Linear(E, self.num_summaries))

    def forward(self, masks, value, need_weights: bool = False):
        """masks [B,K,H,W] in [0,1]; value [B,K,CV,h,w] -> summaries [B,K,Q,E+1] (sums | area), logits or None."""
        h, w = value.shape[-2:]
        m = area_resize(self, masks, (h, w)).unsqueeze(-1)                         # [B,K,h,w,1]
        half = self.num_summaries // 2
        allow = torch.cat([m.expand(-1, -1, -1, -1, half), (1 - m).expand(-1, -1, -1, -1, half)], -1)
        tok = self.input_proj(value.permute(0, 1, 3, 4, 2))
        if self.add_pe:
            tok = tok + self.pos_enc.grid(h, w)
        tok = tok.float()
        feat = self.feature_pred(tok)
        logits = self.weights_pred(tok)
        wgt = logits.sigmoid() * allow                                             # [B,K,h,w,Q]
        sums = torch.einsum('bkhwq,bkhwc->bkqc', wgt, feat)
        area = wgt.flatten(2, 3).sum(2).unsqueeze(-1)
        return torch.cat([sums, area], -1), (logits if need_weights else None)
