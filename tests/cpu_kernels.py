"""TEST-ONLY CPU emulation of cutie_b200.kernels, built on oracle/ math.

Lets the `-m "not gpu"` suite exercise the HOST logic (arena bookkeeping, bucket/permanence rules,
consolidation schedule, the folded-attention algebra in QueryTransformer, InferenceCore.step control
flow) in a container without a GPU, by monkeypatching the ctypes wrappers with functions that honour
exactly the same contracts.  The product never imports this file; on a GPU box the real kernels run and
are compared against the same oracle in tests/test_gpu_*.py.
"""
import math

import torch

from oracle import memory_math as mm


def _cat_rows(rows):
    return torch.cat(list(rows), dim=1)


def affinity_topk(segments, qk, qe, top_k, usage_acc=None, want_sim=False, seed_idx=None):
    from cutie_b200.kernels import kpad_for, KernelError
    keys = _cat_rows([s.key for s in segments])                 # [B,N,CK]
    shr = _cat_rows([s.shrinkage for s in segments])            # [B,N]
    B, N, CK = keys.shape
    if N < top_k:
        raise KernelError('selected index k out of range')
    sim = mm.similarity_direct(keys.transpose(1, 2), shr.unsqueeze(1), qk, qe, dtype=torch.float64)
    # descending similarity, ties toward the lower index (the kernels' documented rule)
    order = torch.argsort(-sim, dim=1, stable=True)[:, :top_k]          # [B,k,Q]
    vals = torch.gather(sim, 1, order).float()
    e = (vals - vals[:, :1]).exp()
    w = e / e.sum(1, keepdim=True)
    kpad = kpad_for(top_k)
    Q = qk.shape[-1]
    idx_o = torch.full((B, Q, kpad), -1, dtype=torch.int32)
    w_o = torch.zeros(B, Q, kpad)
    idx_o[:, :, :top_k] = order.transpose(1, 2).int()
    w_o[:, :, :top_k] = w.transpose(1, 2)
    sim_o = None
    if want_sim:
        sim_o = torch.zeros(B, Q, kpad)
        sim_o[:, :, :top_k] = vals.transpose(1, 2)
    if usage_acc is not None:
        fx = (w.double() * 2.0 ** 40).to(torch.int64)                   # [B,k,Q]
        for b in range(B):
            usage_acc[b].index_add_(0, order[b].reshape(-1), fx[b].reshape(-1))
    return idx_o, w_o, sim_o


def topk_merge(part_val, part_idx, top_k, n_total, usage_acc=None, want_sim=False):
    from cutie_b200.kernels import kpad_for
    B, parts, Q, kpad = part_val.shape
    v = part_val.permute(0, 2, 1, 3).reshape(B, Q, parts * kpad).double().clone()
    i = part_idx.permute(0, 2, 1, 3).reshape(B, Q, parts * kpad).long()
    v[i < 0] = float('-inf')
    # order by (value desc, index asc)
    order = torch.argsort(i, dim=-1, stable=True)
    v, i = torch.gather(v, -1, order), torch.gather(i, -1, order)
    order = torch.argsort(-v, dim=-1, stable=True)[..., :top_k]
    v, i = torch.gather(v, -1, order), torch.gather(i, -1, order)
    live = torch.isfinite(v)
    e = torch.where(live, (v - v[..., :1]).exp(), torch.zeros_like(v))
    w = (e / e.sum(-1, keepdim=True)).float()
    kp = kpad_for(top_k)
    idx_o = torch.full((B, Q, kp), -1, dtype=torch.int32)
    w_o = torch.zeros(B, Q, kp)
    idx_o[..., :top_k] = torch.where(live, i, torch.full_like(i, -1)).int()
    w_o[..., :top_k] = w
    sim_o = None
    if want_sim:
        sim_o = torch.zeros(B, Q, kp)
        sim_o[..., :top_k] = torch.where(live, v, torch.zeros_like(v)).float()
    if usage_acc is not None:
        fx = (w.double() * 2.0 ** 40).to(torch.int64)
        for b in range(B):
            m = live[b].reshape(-1)
            usage_acc[b].index_add_(0, i[b].reshape(-1)[m], fx[b].reshape(-1)[m])
    return idx_o, w_o, sim_o


def readout_gather(idx, w, segments, out=None):
    K = len(segments[0].values)
    vals = [_cat_rows([s.values[k] for s in segments]) for k in range(K)]   # K x [B,N,CV]
    B, Q, kpad = idx.shape
    CV = vals[0].shape[2]
    res = torch.zeros(B, K, CV, Q)
    safe = idx.clamp(min=0).long()
    for b in range(B):
        for k in range(K):
            g = vals[k][b][safe[b].reshape(-1)].reshape(Q, kpad, CV)         # [Q,kpad,CV]
            res[b, k] = (g * w[b].unsqueeze(-1)).sum(1).t()
    if out is not None:
        out.copy_(res)
        return out
    return res


def usage_commit(use_cnt, life_cnt, usage_acc, acc_offset):
    n = use_cnt.shape[1]
    use_cnt += (usage_acc[:, acc_offset:acc_offset + n].double() * 2.0 ** -40).float()
    life_cnt += 1


def bank_append(src, dst_rows):
    dst_rows.copy_(src.transpose(1, 2))


def bank_export(rows, dst):
    dst.copy_(rows.transpose(1, 2))


def upsample2x_add(g, skip):
    B, K = g.shape[:2]
    up = torch.nn.functional.interpolate(g.flatten(0, 1), scale_factor=2, mode='bilinear', align_corners=False)
    return up.reshape(B, K, *up.shape[1:]) + skip.unsqueeze(1)


def prob_to_mask(prob, lut):
    return lut[torch.argmax(prob, dim=0)]


def bank_key_image(key_arena, shr_arena, phys_begin, n, image, mu=None):
    pass        # the operand image only feeds the tcgen05 filter; the CPU emulation reads the fp32 rows


def bank_gather(segments_rows, index, dst_rows):
    allr = _cat_rows(segments_rows)
    for b in range(index.shape[0]):
        dst_rows[b] = allr[b][index[b]]


def consolidate(segments, proto_key, proto_sel, out_values, out_shrinkage, stats=None):
    keys = _cat_rows([s.key for s in segments]).transpose(1, 2)         # [B,CK,N]
    shr = _cat_rows([s.shrinkage for s in segments]).unsqueeze(1)       # [B,1,N]
    sim = mm.similarity_expanded(keys, shr, proto_key.transpose(1, 2), proto_sel.transpose(1, 2))
    aff = mm.dense_softmax(sim)                                          # [B,N,P]
    if stats is not None:
        mx = sim.max(dim=1)[0]
        stats[0].copy_(mx)
        stats[1].copy_((sim - mx.unsqueeze(1)).exp().sum(dim=1))
    for k, ov in enumerate(out_values):
        v = _cat_rows([s.values[k] for s in segments])                   # [B,N,CV]
        ov.copy_(torch.matmul(aff.transpose(1, 2), v))
    out_shrinkage.copy_(torch.matmul(aff.transpose(1, 2), shr.transpose(1, 2)).squeeze(-1))


def obj_summary_accumulate(acc, new):
    acc += new


def qt_linear(x, weight, bias, *, ln=None, pe=None, summary_norm=False, relu=False, residual=None,
              residual_mod=0, xhat_out=None, out=None):
    xin = x
    if summary_norm:
        xin = x[:, :-1] / (x[:, -1:] + 1e-4)
    if ln is not None:
        xin = torch.nn.functional.layer_norm(xin, (xin.shape[-1],), ln[0], ln[1], 1e-5)
        if xhat_out is not None:
            xhat_out.copy_(xin)
    if pe is not None:
        xin = xin + pe
    y = xin @ weight.t()
    if bias is not None:
        y = y + bias
    if relu:
        y = torch.relu(y)
    if residual is not None:
        if residual_mod:
            y = y + residual[torch.arange(y.shape[0]) % residual_mod]
        else:
            y = y + residual
    if out is not None:
        out.copy_(y)
        return out
    return y


def qt_head_fold(a, weight, *, transpose_w, scale, bias_vec=None, num_heads=8):
    M, E = a.shape
    d = E // num_heads
    wx = weight.t() if transpose_w else weight                # Wx[r, c]
    ah = a.reshape(M, num_heads, d)
    out = scale * torch.einsum('mhd,hdc->mhc', ah, wx.reshape(num_heads, d, E))
    dots = None
    if bias_vec is not None:
        dots = scale * torch.einsum('mhd,hd->mh', ah, bias_vec.reshape(num_heads, d))
    return out.contiguous(), dots


def qt_self_attention(qk, v, num_queries, num_heads=8):
    M, E2 = qk.shape
    E = E2 // 2
    d = E // num_heads
    n = M // num_queries
    q = qk[:, :E].reshape(n, num_queries, num_heads, d).transpose(1, 2)
    k = qk[:, E:].reshape(n, num_queries, num_heads, d).transpose(1, 2)
    vv = v.reshape(n, num_queries, num_heads, d).transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), -1)
    return (p @ vv).transpose(1, 2).reshape(M, E)


def qt_aux_mask(pixel, w, b, B, K):
    BK, E, HW = pixel.shape
    logits = (torch.relu(pixel) * w.view(1, E, 1)).sum(1) + b                 # [BK,HW]
    logits = logits.view(B, K, HW)
    p = logits.sigmoid()
    bg = torch.prod(1 - p, dim=1, keepdim=True)
    allp = torch.cat([bg, p], 1).clamp(1e-7, 1 - 1e-7)
    lg = torch.log(allp / (1 - allp))
    fg = (lg[:, 1:] >= lg.max(1, keepdim=True)[0])
    return logits, fg.to(torch.uint8), fg.reshape(BK, HW).sum(1).int()


def qt_pixel_to_query(qfold, pixel, pixel_pe, fg, fg_count, wv, bv, num_queries, num_heads=8, splits=None):
    M, H, E = qfold.shape
    BK, _, HW = pixel.shape
    Q = num_queries
    d = E // H
    kin = pixel + pixel_pe                                                    # [BK,E,HW]
    qf = qfold.reshape(BK, Q, H, E)
    s = torch.einsum('nqhc,ncp->nqhp', qf, kin)                               # [BK,Q,H,HW]
    f = fg.reshape(BK, HW).bool()
    half = Q // 2
    blocked = torch.cat([(~f)[:, None].expand(BK, half, HW), f[:, None].expand(BK, half, HW)], 1).clone()
    none_fg = (fg_count == 0)
    all_fg = (fg_count == HW)
    blocked[none_fg, :half] = False
    blocked[all_fg, half:] = False
    s = s.masked_fill(blocked[:, :, None, :], float('-inf'))
    p = torch.softmax(s, -1)
    z = torch.einsum('nqhp,ncp->nqhc', p, pixel)                              # [BK,Q,H,E]
    attn = torch.einsum('nqhc,hdc->nqhd', z, wv.reshape(H, d, E)) + bv.reshape(H, d)
    return attn.reshape(M, E)


def qt_pixel_to_query_tiles(qfold, pixel, pixel_pe, fg, fg_count, num_queries, num_heads=8):
    # the emulation has no tile workspace: hand the inputs to the chain's combine op
    return (qfold, pixel, pixel_pe, fg, fg_count), (pixel.shape[2] + 63) // 64


def qt_chain_run(chain):
    """Emulation of cutie_qt_chain: the recorded ops one after the other (phases only constrain ordering)."""
    g = globals()
    for name, args, kw, outs, _phase in chain.ops:
        if name == 'qt_linear':
            g[name](*args, out=outs[0], **kw)
        elif name == 'qt_head_fold':
            o, d = g[name](*args, **kw)
            outs[0].copy_(o)
            if outs[1] is not None:
                outs[1].copy_(d)
        elif name == 'qt_self_attention':
            outs[0].copy_(g[name](*args))
        else:
            ws, _tiles, wv, bv, _BK, Q, H = args
            outs[0].copy_(qt_pixel_to_query(*ws, wv, bv, Q, H))


def qt_query_to_pixel(kfold, kdots, vfold, out_bias, pixel, pixel_pe, num_queries, num_heads=8, out=None):
    BK, E, HW = pixel.shape
    Q, H = num_queries, num_heads
    kf = kfold.reshape(BK, Q, H, E)
    s = torch.einsum('ncp,nqhc->nphq', pixel + pixel_pe, kf) + kdots.reshape(BK, Q, H).permute(0, 2, 1)[:, None]
    p = torch.softmax(s, -1)                                                  # over the Q queries
    upd = torch.einsum('nphq,nqhc->ncp', p, vfold.reshape(BK, Q, H, E))
    res = pixel + upd + out_bias.view(1, E, 1)
    if out is not None:
        out.copy_(res)
        return out
    return res


def bias_act_(y, bias, z=None, relu=False):
    assert y.dim() == 4 and (y.is_contiguous() or y.is_contiguous(memory_format=torch.channels_last))
    y.add_(bias.detach().view(1, -1, 1, 1))
    if z is not None:
        y.add_(z)
    return torch.relu_(y) if relu else y


def bias_relu_maxpool(y, bias):
    import torch.nn.functional as F
    return F.max_pool2d(torch.relu(y + bias.detach().view(1, -1, 1, 1)), 3, stride=2, padding=1)


def segment_tail(x):
    import torch.nn.functional as F
    from cutie_b200.utils.tensor_utils import aggregate
    logits = F.interpolate(aggregate(torch.sigmoid(x), dim=1), scale_factor=4, mode='bilinear', align_corners=False)
    return logits, F.softmax(logits, dim=1)


def conv3x3_c1(x, weight, bias, relu_input=False):
    import torch.nn.functional as F
    return F.conv2d(torch.relu(x) if relu_input else x, weight, bias, padding=1)


def conv_weight_image(weight):
    return weight.detach()                      # the emulation's "operand image" is the weight itself


def conv_tc(x, weight_image, bias, cout, ksize=3, stride=1, residual=None, relu_in=False, relu_out=False, units_per_cta=None, counters=None):
    import torch.nn.functional as F
    assert weight_image.shape[0] == cout and weight_image.shape[2] == ksize
    y = F.conv2d(torch.relu(x) if relu_in else x, weight_image, bias, stride=stride, padding=ksize // 2)
    if residual is not None:
        y = y + residual
    return torch.relu(y) if relu_out else y


def area_pool(x, f):
    import torch.nn.functional as F
    lead = x.shape[:-2]
    y = F.avg_pool2d(x.reshape(-1, 1, *x.shape[-2:]), f)
    return y.reshape(*lead, *y.shape[-2:])


def eca_scale_add_(y, x, conv1d_weight):
    import torch.nn.functional as F
    k = conv1d_weight.shape[-1]
    gate = torch.sigmoid(F.conv1d(y.mean(dim=(2, 3)).unsqueeze(1), conv1d_weight, padding=(k - 1) // 2))
    return y.mul_(gate.transpose(1, 2).unsqueeze(-1)).add_(x)


def gated_update(h, v):
    d = h.shape[2]
    f, u, n = torch.sigmoid(v[:, :, :d]), torch.sigmoid(v[:, :, d:2 * d]), torch.tanh(v[:, :, 2 * d:])
    return f * h * (1 - u) + u * n


ALL = ['last_candidate_counts', 'bias_act_', 'bias_relu_maxpool', 'segment_tail', 'conv3x3_c1', 'conv_weight_image', 'conv_tc', 'area_pool', 'eca_scale_add_', 'gated_update', 'affinity_topk', 'topk_merge', 'readout_gather', 'usage_commit', 'bank_append', 'bank_export', 'bank_gather', 'bank_key_image', 'upsample2x_add', 'prob_to_mask',
       'consolidate', 'obj_summary_accumulate', 'qt_linear', 'qt_head_fold', 'qt_self_attention',
       'qt_aux_mask', 'qt_pixel_to_query', 'qt_query_to_pixel', 'qt_pixel_to_query_tiles', 'qt_chain_run']


def last_candidate_counts():
    return None


def install(monkeypatch=None):
    """Swap the ctypes wrappers for the emulations above (test-only)."""
    import cutie_b200.kernels as K_
    g = globals()
    for name in ALL:
        if monkeypatch is not None:
            monkeypatch.setattr(K_, name, g[name])
        else:
            setattr(K_, name, g[name])
