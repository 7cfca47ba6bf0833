"""The HOT plan inside the runtime: InferenceCore.step with the tcgen05 candidate filter over the key operand image
(set_tc_min_tokens(256): every read goes through csrc/affinity_f16.cu), on banks that wrap the ring arena, evict by FIFO,
consolidate into long-term memory and compact it -- i.e. every path that keeps the image in step with the arena.

  * free-running streams on the tcgen05 plan vs the SAME streams on the exact-scan plan: the two plans are bit-identical
    in selection and weights, so the whole stream must be bit-identical (torch.equal on the logits of every frame);
  * one teacher-forced frame vs the CPU oracle from the live state of a pre-filled >= 8k-token bank after the ring wrapped;
  * top-k parity on NETWORK-DERIVED, temporally coherent features (SURVEY.md Appendix B take-away 4): keys / shrinkage /
    selection of the key projection over a coherent clip, >= 8k tokens, against the float64 ground truth and the exact scan.

(The other end-to-end files keep their banks below the plan threshold: they cover the exact-scan plan.)"""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method='thread')]


def _net(cfg):
    from cutie_b200.model.cutie import CUTIE
    from oracle.synth import synthetic_state_dict
    net = CUTIE(cfg).eval()
    net.load_state_dict(synthetic_state_dict(net.state_dict(), 0))
    return net.cuda()


@pytest.fixture
def K_():
    import cutie_b200.kernels as K
    yield K
    K.set_tc_min_tokens(-1)


def _run_stream(K_, cfg, net, frames, mask, objects, graphs, prefill=None):
    from cutie_b200.inference.inference_core import InferenceCore
    proc = InferenceCore(net, cfg=cfg, use_cuda_graphs=graphs)
    logits, sizes = [], []
    before = K_.image_level_launches()
    with torch.inference_mode():
        for ti in range(frames.shape[0]):
            if ti == 0:
                proc.step(frames[0].cuda(), mask.cuda(), objects=objects)
                if prefill is not None:
                    key, shr, vals = prefill
                    proc.memory.work_mem.add(key.cuda(), {o: vals[:, i].cuda() for i, o in enumerate(objects)}, shr.cuda(),
                                             None, as_permanent='no')
            else:
                proc.step(frames[ti].cuda())
                logits.append(proc.last_logits.clone())
            m = proc.memory
            sizes.append((m.work_mem.size(0), m.long_mem.size(0) if m.use_long_term else 0))
    return logits, sizes, K_.image_level_launches() - before, proc


LT = dict(max_mem_frames=6, min_mem_frames=3, num_prototypes=64, max_num_tokens=300, buffer_tokens=100)


@pytest.mark.parametrize('name,over,T,graphs', [
    ('fifo ring wrap + eviction', dict(mem_every=1, max_mem_frames=6), 24, False),
    ('fifo ring wrap + eviction, CUDA graphs', dict(mem_every=2, max_mem_frames=5), 24, True),
    ('long-term: consolidation, prototypes, obsolete-feature compaction', dict(mem_every=1, use_long_term=True, long_term=LT), 26,
     False),
])
def test_stream_on_the_tcgen05_plan_is_bit_identical_to_the_exact_scan_plan(K_, name, over, T, graphs):
    from cutie_b200.config import default_config
    from oracle.synth import synthetic_video
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = default_config(**over)
    net = _net(cfg)
    frames, mask = synthetic_video(T, 240, 432, 3, seed=6)               # 15 x 27 = 405 tokens per frame
    K_.set_tc_min_tokens(256)
    hot, sizes_hot, launches, proc = _run_stream(K_, cfg, net, frames, mask, [1, 2, 3], graphs)
    assert launches >= T - 4, f'{name}: only {launches} reads went through the image plan'
    K_.set_tc_min_tokens(1 << 40)
    cold, sizes_cold, launches_cold, _ = _run_stream(K_, cfg, net, frames, mask, [1, 2, 3], graphs)
    assert launches_cold == 0 and sizes_hot == sizes_cold
    # the bank went through its whole life cycle
    assert max(s[0] for s in sizes_hot) > min(s[0] for s in sizes_hot[3:]), 'no eviction happened'
    if over.get('use_long_term'):
        longs = [s[1] for s in sizes_hot]
        assert max(longs) > 0 and any(b < a for a, b in zip(longs, longs[1:])), 'long-term memory was never compacted'
    for ti, (a, b) in enumerate(zip(hot, cold)):
        assert torch.equal(a, b), f'{name}: frame {ti + 1} differs by {float((a - b).abs().max()):.3e}'


@pytest.mark.parametrize('optimised', [False, True])
def test_teacher_forced_frame_from_a_prefilled_ring_after_the_wrap(K_, optimised):
    """>= 8k pre-filled tokens + real memory frames on top until the ring has wrapped and evicted, then ONE frame against
    the CPU oracle from the exported live state (near-tied top-k members / foreground pixels reconciled vs float64).
    optimised: the product configuration -- optimize_for_inference(): BN folded, 3x3 / 1x1 convolutions on cutie_conv_tc."""
    from cutie_b200.config import default_config
    from oracle.cpu_core import OracleCore
    from oracle.state_sync import ForegroundReconciler, SelectionReconciler, export_state_to_oracle
    from oracle.synth import synthetic_state_dict, synthetic_video
    from cutie_b200.model.cutie import CUTIE
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    HW, frames_in_bank = 405, 24
    cfg = default_config(mem_every=2, max_mem_frames=frames_in_bank)
    net = _net(cfg)
    if optimised:
        net.optimize_for_inference()
    T = 15
    frames, mask = synthetic_video(T + 1, 240, 432, 3, seed=9)
    g = torch.Generator().manual_seed(1)
    n = (frames_in_bank - 3) * HW                                         # 8 505 tokens: a few frames below the FIFO limit
    prefill = (torch.randn(1, 64, n, generator=g), 1 + torch.randn(1, 1, n, generator=g) ** 2, torch.randn(1, 3, 256, n, generator=g))
    K_.set_tc_min_tokens(256)
    _, sizes, launches, proc = _run_stream(K_, cfg, net, frames[:T], mask, [1, 2, 3], False, prefill)
    assert launches >= T - 2 and sizes[-1][0] >= 8000
    # 7 memory frames went on top of the pre-fill: more than the ring holds, so its head advanced (FIFO) and it wrapped
    added = n + 405 * sum(1 for ti in range(1, T) if ti % 2 == 0)
    assert added > sizes[-1][0] - 405 and sizes[-1][0] == sizes[-3][0], (added, sizes[-4:])
    cpu_net = CUTIE(cfg).eval()
    cpu_net.load_state_dict(synthetic_state_dict(cpu_net.state_dict(), 0))
    oc = export_state_to_oracle(proc, OracleCore(cpu_net, cfg))
    rec, fgr = SelectionReconciler(cfg.top_k), ForegroundReconciler()
    oc.selection_hook, oc.fg_hook = rec, fgr
    orig, orig_aux = K_.affinity_topk, K_.qt_aux_mask

    def spy(*a, **k):
        out = orig(*a, **k)
        rec.gpu_idx = out[0].clone()
        return out

    def spy_aux(*a, **k):
        out = orig_aux(*a, **k)
        fgr.gpu_fg.append(out[1].clone())
        return out
    K_.affinity_topk, K_.qt_aux_mask = spy, spy_aux
    try:
        with torch.inference_mode():
            proc.step(frames[T].cuda())
            oc.step(frames[T])
    finally:
        K_.affinity_topk, K_.qt_aux_mask = orig, orig_aux
    d = float((proc.last_logits.cpu() - oc.last_logits).abs().max())
    if optimised:
        rep = net.conv_epilogues.report()['layers']
        assert rep.get('tc', 0) >= 80, rep                                  # the tensor-core convolutions really ran
    print(f'teacher-forced frame on the hot plan (optimised={optimised}, {proc.memory.work_mem.size(0)} tokens): max |logit diff| = {d:.3e}, '
          f'{rec.flips} near-tied selections, {fgr.flips} near-tied foreground pixels')
    assert d < 1e-3, d


def test_topk_parity_on_network_derived_temporally_coherent_features(K_):
    """Keys / shrinkage of 21 consecutive frames of a coherent clip through the network's own encoder + key projection
    (8 505 memory tokens that are near-duplicates of each other frame to frame), queries = the next frame: the hard
    distribution of SURVEY.md Appendix B -- top-30 sets vs float64 truth, and bit-identity with the exact scan."""
    from cutie_b200.config import default_config
    from oracle import memory_math as mm
    from oracle.synth import synthetic_video
    torch.backends.cudnn.allow_tf32 = False
    cfg = default_config()
    net = _net(cfg)
    frames, _ = synthetic_video(22, 240, 432, 1, seed=12)
    keys, shrs = [], []
    with torch.inference_mode():
        for ti in range(22):
            img = frames[ti].cuda()[None]
            pad_w = (16 - img.shape[-1] % 16) % 16
            img = torch.nn.functional.pad(img, (pad_w // 2, pad_w - pad_w // 2, 0, 0))
            ms, _ = net.encode_image(img)
            k, s, e = net.transform_key(ms[0])
            if ti < 21:
                keys.append(k.flatten(2)), shrs.append(s.flatten(2))
            else:
                qk, qe = k.flatten(2).contiguous(), e.flatten(2).contiguous()
    key = torch.cat(keys, 2).transpose(1, 2).contiguous()                 # [1, N, 64] token-major
    shr = torch.cat(shrs, 2)[:, 0].contiguous()                           # [1, N]
    N, Q, top_k = key.shape[1], qk.shape[2], 30
    assert N >= 8000
    img_t = torch.zeros(1, K_.key_image_tiles(N), K_.KEY_IMAGE_FLOATS, device='cuda')
    K_.bank_key_image(key, shr, 0, N, img_t)
    K_.set_tc_min_tokens(256)
    before = K_.image_level_launches()
    idx, w, sim = K_.affinity_topk([K_.BankSegment(key, shr, (), img_t, 0)], qk, qe, top_k, want_sim=True)
    assert K_.image_level_launches() == before + 1
    K_.set_tc_min_tokens(1 << 40)
    idx_x, w_x, sim_x = K_.affinity_topk([K_.BankSegment(key, shr, ())], qk, qe, top_k, want_sim=True)
    assert torch.equal(idx, idx_x) and torch.equal(w, w_x) and torch.equal(sim, sim_x)
    truth = mm.similarity_direct(key.cpu().transpose(1, 2), shr.cpu().unsqueeze(1), qk.cpu(), qe.cpu(), dtype=torch.float64)
    n_dec, n_dec_eq, n_eq, n_all = mm.topk_set_agreement(idx[:, :, :top_k].cpu().long().transpose(1, 2), truth, top_k,
                                                         rel_noise=1e-5)
    gap = truth.topk(top_k + 1, dim=1)[0]
    print(f'network-derived features: S in [{float(truth.min()):.1f}, {float(truth.max()):.2f}], 30th/31st gap min '
          f'{float((gap[:, -2] - gap[:, -1]).min()):.2e}; {n_dec} of {n_all} queries decidable in fp32, {n_eq} sets identical to float64')
    assert n_dec > 0 and n_dec_eq == n_dec
