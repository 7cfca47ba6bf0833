"""CPU oracle for the object transformer as pure functions of a
state_dict.  TEST INFRASTRUCTURE ONLY (see oracle/memory_math.py header).

Restates, with explicit matmuls instead of nn.MultiheadAttention:
  cutie/model/transformer/object_transformer.py:114-205  (QueryTransformer.forward, _get_aux_mask)
  cutie/model/transformer/object_transformer.py:35-73    (QueryTransformerBlock.forward)
  cutie/model/transformer/transformer_layers.py:12-136   (SelfAttention, CrossAttention, FFN, PixelFFN)
  cutie/model/channel_attn.py:7-39                       (CAResBlock)
  cutie/model/transformer/positional_encoding.py:20-97   (PositionalEncoding)
  cutie/utils/tensor_utils.py:47-54                      (aggregate)
Weights are looked up by the reference's state_dict key names under a prefix.
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def sinusoid_pe(h: int, w: int, embed_dim: int = 256, scale: float = 32.0, temperature: float = 128.0,
                dtype=torch.float32) -> torch.Tensor:
    """positional_encoding.py:29-31,72-85 -> [h,w,embed_dim]; x-half then y-half, sin/cos interleaved."""
    half = int(math.ceil(embed_dim / 4) * 2)
    inv_freq = 1.0 / (temperature ** (torch.arange(0, half, 2, dtype=torch.float32) / half))
    ys = torch.arange(h, dtype=torch.float32)
    xs = torch.arange(w, dtype=torch.float32)
    ys = ys / (ys[-1] + 1e-6) * scale
    xs = xs / (xs[-1] + 1e-6) * scale
    ay = ys[:, None] * inv_freq[None]
    ax = xs[:, None] * inv_freq[None]
    ey = torch.stack((ay.sin(), ay.cos()), dim=-1).flatten(-2)   # [h,half]
    ex = torch.stack((ax.sin(), ax.cos()), dim=-1).flatten(-2)   # [w,half]
    pe = torch.zeros(h, w, 2 * half, dtype=torch.float32)
    pe[:, :, :half] = ex[None]
    pe[:, :, half:] = ey[:, None]
    return pe.to(dtype)


def aggregate_logits(prob: torch.Tensor, dim: int) -> torch.Tensor:
    """tensor_utils.py:47-54: prepend bg = prod(1-p), clamp to [1e-7, 1-1e-7], log-odds."""
    bg = torch.prod(1 - prob, dim=dim, keepdim=True)
    p = torch.cat([bg, prob], dim).clamp(1e-7, 1 - 1e-7)
    return torch.log(p / (1 - p))


def foreground_map(aux_logits: torch.Tensor) -> torch.Tensor:
    """object_transformer.py:185-192: fg[b,k,p] = (logit_k >= max over {bg, 1..K}).  [B,K,h,w] -> bool [B,K,HW]."""
    lg = aggregate_logits(aux_logits.sigmoid(), dim=1)
    return (lg[:, 1:] >= lg.max(dim=1, keepdim=True)[0]).flatten(2)


def attention_block_mask(fg: torch.Tensor, num_queries: int) -> torch.Tensor:
    """object_transformer.py:193-205.  fg bool [B,K,HW] -> blocked bool [B*K, num_queries, HW]
    (identical for all heads): first half of the queries may only look at foreground pixels, second
    half only at background; a row that would be fully blocked is fully opened (:203)."""
    B, K, HW = fg.shape
    half = num_queries // 2
    blocked = torch.cat([(~fg)[:, :, None].expand(B, K, half, HW),
                         fg[:, :, None].expand(B, K, half, HW)], dim=2).reshape(B * K, num_queries, HW).clone()
    full = blocked.all(dim=-1)
    blocked[full] = False
    return blocked


def _ln(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def _mha(q_in, k_in, v_in, in_w, in_b, out_w, out_b, num_heads, blocked=None):
    """nn.MultiheadAttention (batch_first, packed in_proj) spelled out.
    q_in [N,L,E], k_in/v_in [N,S,E], blocked bool [N,L,S] or None.  Returns ([N,L,E], P [N,H,L,S])."""
    E = q_in.shape[-1]
    d = E // num_heads
    wq, wk, wv = in_w[:E], in_w[E:2 * E], in_w[2 * E:]
    bq, bk, bv = in_b[:E], in_b[E:2 * E], in_b[2 * E:]
    N, L, _ = q_in.shape
    S = k_in.shape[1]
    q = (q_in @ wq.t() + bq).reshape(N, L, num_heads, d).transpose(1, 2)
    k = (k_in @ wk.t() + bk).reshape(N, S, num_heads, d).transpose(1, 2)
    v = (v_in @ wv.t() + bv).reshape(N, S, num_heads, d).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    if blocked is not None:
        s = s.masked_fill(blocked[:, None], float('-inf'))
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(N, L, E)
    return o @ out_w.t() + out_b, p


def ca_res_block(x: torch.Tensor, sd: Dict[str, torch.Tensor], pre: str) -> torch.Tensor:
    """channel_attn.py:26-39 (in_dim == out_dim, residual)."""
    r = x
    x = F.conv2d(F.relu(x), sd[pre + 'conv1.weight'], sd[pre + 'conv1.bias'], padding=1)
    x = F.conv2d(F.relu(x), sd[pre + 'conv2.weight'], sd[pre + 'conv2.bias'], padding=1)
    b, c = x.shape[:2]
    kw = sd[pre + 'conv.weight']
    wgt = F.conv1d(x.mean(dim=(2, 3)).view(b, 1, c), kw, padding=(kw.shape[-1] - 1) // 2)
    wgt = wgt.transpose(-1, -2).unsqueeze(-1).sigmoid()
    if (pre + 'downsample.weight') in sd:
        r = F.conv2d(r, sd[pre + 'downsample.weight'], sd[pre + 'downsample.bias'])
    return x * wgt + r


def query_transformer(pixel: torch.Tensor, obj_summaries: torch.Tensor, sd: Dict[str, torch.Tensor],
                      prefix: str = 'object_transformer.', num_heads: int = 8, num_blocks: int = 3,
                      pe_scale: float = 32.0, pe_temperature: float = 128.0,
                      trace: Optional[dict] = None, fg_hook=None):
    """QueryTransformer.forward (object_transformer.py:114-177), inference mode, selector=None.

    pixel [B,K,E,h,w]; obj_summaries [B,K,T,Q,E+1].  Returns (pixel_out [B,K,E,h,w], aux_logits list).
    If `trace` is a dict it receives intermediate tensors (used by the module-level parity tests).
    `fg_hook(stage, aux_logits [B,K,h,w], fg bool [B,K,HW]) -> fg` (test-only) lets a checker substitute an equally valid
    foreground map on near-tied pixels (oracle/state_sync.ForegroundReconciler), like OracleCore.selection_hook does
    for near-tied top-k members.
    """
    g = lambda n: sd[prefix + n]
    B, K, E, h, w = pixel.shape
    HW = h * w
    Q = g('query_init.weight').shape[0]
    T = obj_summaries.shape[2]
    osum = obj_summaries.reshape(B * K, T, Q, E + 1)
    vals = osum[..., :-1].sum(1) / (osum[..., -1:].sum(1) + 1e-4)                       # :126-132
    x = g('query_init.weight')[None] + vals @ g('summary_to_query_init.weight').t() + g('summary_to_query_init.bias')
    qpe = g('query_emb.weight')[None] + vals @ g('summary_to_query_emb.weight').t() + g('summary_to_query_emb.bias')

    flat = pixel.flatten(0, 1)
    p_init = F.conv2d(flat, g('pixel_init_proj.weight'), g('pixel_init_proj.bias'))     # :141
    p_emb = F.conv2d(flat, g('pixel_emb_proj.weight'), g('pixel_emb_proj.bias'))        # :142
    pe = sinusoid_pe(h, w, E, pe_scale, pe_temperature, pixel.dtype).reshape(1, HW, E)
    pixel_pe = pe + p_emb.flatten(2).transpose(1, 2)                                    # :143-145  [BK,HW,E]
    pix = p_init                                                                        # [BK,E,h,w]

    def mask_pred(i, t):
        return F.conv2d(F.relu(t), g(f'mask_pred.{i}.1.weight'), g(f'mask_pred.{i}.1.bias')).reshape(B, K, h, w)

    def fg_of(stage, lg):
        fg = foreground_map(lg)
        return fg if fg_hook is None else fg_hook(stage, lg, fg)

    logits: List[torch.Tensor] = [mask_pred(0, pix)]
    blocked = attention_block_mask(fg_of(0, logits[0]), Q)
    if trace is not None:
        trace.update(query0=x.clone(), query_pe=qpe.clone(), pixel_pe=pixel_pe.clone(), pixel0=pix.clone(),
                     blocked0=blocked.clone())
    for i in range(num_blocks):
        bp = f'blocks.{i}.'
        pf = pix.flatten(2).transpose(1, 2)                                             # [BK,HW,E]
        # read_from_pixel: transformer_layers.py:66-98 (norm, +pe on q/k, residual = normed x)
        xn = _ln(x, g(bp + 'read_from_pixel.norm.weight'), g(bp + 'read_from_pixel.norm.bias'))
        a, _ = _mha(xn + qpe, pf + pixel_pe, pf, g(bp + 'read_from_pixel.cross_attn.in_proj_weight'),
                    g(bp + 'read_from_pixel.cross_attn.in_proj_bias'),
                    g(bp + 'read_from_pixel.cross_attn.out_proj.weight'),
                    g(bp + 'read_from_pixel.cross_attn.out_proj.bias'), num_heads, blocked)
        x = xn + a
        if trace is not None:
            trace[f'b{i}_after_rfp'] = x.clone()
        # self attention: transformer_layers.py:27-41
        xn = _ln(x, g(bp + 'self_attn.norm.weight'), g(bp + 'self_attn.norm.bias'))
        a, _ = _mha(xn + qpe, xn + qpe, xn, g(bp + 'self_attn.self_attn.in_proj_weight'),
                    g(bp + 'self_attn.self_attn.in_proj_bias'), g(bp + 'self_attn.self_attn.out_proj.weight'),
                    g(bp + 'self_attn.self_attn.out_proj.bias'), num_heads)
        x = xn + a
        # ffn: transformer_layers.py:113-118 (residual is the un-normed x)
        hdn = F.relu(_ln(x, g(bp + 'ffn.norm.weight'), g(bp + 'ffn.norm.bias')) @ g(bp + 'ffn.linear1.weight').t()
                     + g(bp + 'ffn.linear1.bias'))
        x = x + hdn @ g(bp + 'ffn.linear2.weight').t() + g(bp + 'ffn.linear2.bias')
        if trace is not None:
            trace[f'b{i}_query'] = x.clone()
        # read_from_query: no norm (object_transformer.py:29-32), softmax over the Q queries
        a, _ = _mha(pf + pixel_pe, x + qpe, x, g(bp + 'read_from_query.cross_attn.in_proj_weight'),
                    g(bp + 'read_from_query.cross_attn.in_proj_bias'),
                    g(bp + 'read_from_query.cross_attn.out_proj.weight'),
                    g(bp + 'read_from_query.cross_attn.out_proj.bias'), num_heads)
        pf = pf + a
        if trace is not None:
            trace[f'b{i}_pixel_flat'] = pf.clone()
        # pixel_ffn: transformer_layers.py:127-136
        pix = ca_res_block(pf.transpose(1, 2).reshape(B * K, E, h, w), sd, prefix + bp + 'pixel_ffn.conv.')
        logits.append(mask_pred(i + 1, pix))                                            # :164-167
        blocked = attention_block_mask(fg_of(i + 1, logits[-1]), Q)
        if trace is not None:
            trace[f'b{i}_pixel'] = pix.clone()
    return pix.reshape(B, K, E, h, w), logits
