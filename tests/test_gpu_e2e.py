"""GPU end-to-end parity: InferenceCore.step on the fused CUDA path vs (a) the committed reference
fixtures, free-running, and (b) the CPU oracle's full-frame restatement at a larger size, both
free-running and teacher-forced.  Bar: <= 1e-3 max-abs on segmentation logits (BASELINE.json)."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu

LT_SMALL = dict(max_mem_frames=4, min_mem_frames=2, num_prototypes=16, max_num_tokens=60, buffer_tokens=20)


def _net(cfg, cuda=True):
    from cutie_b200.model.cutie import CUTIE
    from oracle.synth import synthetic_state_dict
    net = CUTIE(cfg).eval()
    net.load_state_dict(synthetic_state_dict(net.state_dict(), 0))
    return net.cuda() if cuda else net


def _sizes(proc):
    m = proc.memory
    row = []
    for b in sorted(m.work_mem.buckets):
        row += [b, m.work_mem.size(b), m.work_mem.perm_size(b), m.long_mem.size(b) if m.use_long_term else 0]
    return row


@pytest.mark.parametrize('name,over,T,K', [
    ('fifo', dict(mem_every=2, max_mem_frames=3), 10, 3),
    ('longterm', dict(mem_every=1, use_long_term=True, long_term=LT_SMALL), 14, 2),
])
def test_free_running_vs_reference_fixture(name, over, T, K):
    from cutie_b200.config import default_config
    from cutie_b200.inference.inference_core import InferenceCore
    from oracle.synth import synthetic_video
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(GOLDEN, f'e2e_{name}.npz'))
    cfg = default_config(**over)
    proc = InferenceCore(_net(cfg), cfg=cfg)
    frames, mask = synthetic_video(T, 96, 160, K, seed=3)
    li, worst = 0, 0.0
    with torch.inference_mode():
        for ti in range(T):
            if ti == 0:
                prob = proc.step(frames[0].cuda(), mask.cuda(), objects=list(range(1, K + 1)))
            else:
                prob = proc.step(frames[ti].cuda())
            assert _sizes(proc) == [int(x) for x in g['sizes'][ti] if x >= 0]
            if ti > 0:
                worst = max(worst, float(np.abs(proc.last_logits.cpu().numpy() - g['logits'][li:li + 1]).max()))
                li += 1
    assert worst < 1e-3, worst
    assert float((prob.cpu() - torch.from_numpy(g['final_prob'])).abs().max()) < 1e-3


def test_buckets_and_delete_vs_reference_fixture():
    """Second object introduced later -> second bucket; then an object is deleted."""
    from cutie_b200.config import default_config
    from cutie_b200.inference.inference_core import InferenceCore
    from oracle.synth import synthetic_video
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(GOLDEN, 'e2e_buckets.npz'))
    cfg = default_config(mem_every=2, max_mem_frames=3)
    proc = InferenceCore(_net(cfg), cfg=cfg)
    frames, _ = synthetic_video(8, 96, 160, 3, seed=3)
    first, second = torch.from_numpy(g['first_mask']).cuda(), torch.from_numpy(g['second_mask']).cuda()
    with torch.inference_mode():
        for ti in range(8):
            if ti == 0:
                prob = proc.step(frames[0].cuda(), first, objects=[1, 2])
            elif ti == 3:
                prob = proc.step(frames[3].cuda(), second, objects=[7])
            elif ti == 6:
                proc.delete_objects([1])
                prob = proc.step(frames[6].cuda())
            else:
                prob = proc.step(frames[ti].cuda())
            if ti == 4:
                assert float(np.abs(proc.last_logits.cpu().numpy() - g['logits_f4']).max()) < 1e-3
                assert len(proc.memory.work_mem.buckets) == 2
    assert float(np.abs(proc.last_logits.cpu().numpy() - g['logits']).max()) < 1e-3
    assert float((prob.cpu() - torch.from_numpy(g['final_prob'])).abs().max()) < 1e-3


@pytest.mark.parametrize('over,H,W,K,T', [
    (dict(mem_every=3, max_mem_frames=4), 240, 432, 3, 12),
    (dict(mem_every=2, use_long_term=True, long_term=dict(max_mem_frames=4, min_mem_frames=2, num_prototypes=64,
                                                          max_num_tokens=600, buffer_tokens=100)), 240, 432, 3, 16),
    (dict(mem_every=3, max_mem_frames=4, flip_aug=True, chunk_size=1), 240, 432, 3, 8),
    (dict(mem_every=2, max_mem_frames=3, top_k=50), 480, 854, 3, 5),          # 480p: 30x54 = 1620 tokens/frame
    (dict(mem_every=2, max_mem_frames=3), 480, 854, 3, 5),
])
def test_teacher_forced_vs_cpu_oracle(over, H, W, K, T):
    """Sizes the fixtures do not cover.  Every frame starts from the CPU oracle's exact state (a random-weight
    recurrent net amplifies 1e-5 differences chaotically when free-running), runs ONE step on each side and
    compares the segmentation logits, the new memory tokens and the sensory state.  Near-tie top-k
    disagreements are arbitrated against float64 ground truth (tests/state_sync.SelectionReconciler)."""
    import cutie_b200.kernels as K_
    from cutie_b200.config import default_config
    from cutie_b200.inference.inference_core import InferenceCore
    from oracle.cpu_core import OracleCore
    from oracle.synth import synthetic_video
    from tests.state_sync import SelectionReconciler, load_state_from_oracle
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = default_config(**over)
    proc, oracle = InferenceCore(_net(cfg), cfg=cfg), OracleCore(_net(cfg, cuda=False), cfg)
    frames, mask = synthetic_video(T, H, W, K, seed=5)
    objs = list(range(1, K + 1))
    rec = SelectionReconciler(cfg.top_k)
    oracle.selection_hook = rec
    real_topk = K_.affinity_topk

    def spy(*a, **kw):
        r = real_topk(*a, **kw)
        rec.gpu_idx = r[0]
        return r

    def one_frame(ti):
        load_state_from_oracle(proc, oracle, 'cuda')
        if ti == 0:
            pg = proc.step(frames[0].cuda(), mask.cuda(), objects=objs)
            pc = oracle.step(frames[0], mask, objects=objs)
            err = 0.0
        else:
            pg = proc.step(frames[ti].cuda())
            pc = oracle.step(frames[ti])
            err = float((proc.last_logits.cpu() - oracle.last_logits).abs().max())
        assert err < 1e-3, (ti, err)
        assert float((pg.cpu() - pc).abs().max()) < 1e-3
        m = proc.memory
        assert m.work_mem.size(0) == oracle.work.size(0)
        assert torch.allclose(m.work_mem.key[0].cpu(), oracle.work.k[0], atol=1e-4)
        assert torch.allclose(m.work_mem.value[objs[-1]].cpu(), oracle.work.v[objs[-1]], rtol=1e-3, atol=2e-3)
        if cfg.use_long_term:
            assert m.long_mem.size(0) == oracle.long.size(0)
            if 0 in oracle.work.use:            # no temporary tokens (hence no counters) before the second memory frame
                assert torch.allclose(m.work_mem.use_cnt[0].cpu(), oracle.work.use[0], atol=1e-4)
            if m.long_mem.size(0):       # prototype order may differ where two usages tie to 1e-7
                assert torch.allclose(m.long_mem.shrinkage[0].cpu().sort(-1)[0], oracle.long.s[0].sort(-1)[0],
                                      rtol=1e-3, atol=1e-3)
        for o in objs:
            assert float((m.sensory[o].cpu() - oracle.sensory[o]).abs().max()) < 2e-3
        return err

    K_.affinity_topk = spy
    try:
        with torch.inference_mode():
            worst = max(one_frame(ti) for ti in range(T))
    finally:
        K_.affinity_topk = real_topk
    assert worst < 1e-3
    assert rec.flips <= 0.02 * max(rec.queries, 1) + 2


def test_cuda_graph_frame_path_matches_eager():
    """use_cuda_graphs=True replays the encoder and the fusion/transformer/decoder regions as CUDA graphs; the
    results must equal the eager path (same kernels, same order) over propagated and memory frames."""
    from cutie_b200.config import default_config
    from cutie_b200.inference.inference_core import InferenceCore
    import cutie_b200.kernels as K_
    from oracle.synth import synthetic_video
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    cfg = default_config(mem_every=3, max_mem_frames=3)
    net = _net(cfg)
    eager, graphed = InferenceCore(net, cfg=cfg), InferenceCore(net, cfg=cfg, use_cuda_graphs=True)
    frames, mask = synthetic_video(9, 240, 432, 3, seed=7)
    with torch.inference_mode():
        for ti in range(9):
            a = frames[ti].cuda()
            if ti == 0:
                pe = eager.step(a, mask.cuda(), objects=[1, 2, 3])
                pg = graphed.step(a, mask.cuda(), objects=[1, 2, 3])
            else:
                n0 = K_.LAUNCH_COUNT
                pe = eager.step(a)
                n_eager = K_.LAUNCH_COUNT - n0
                ncap = len(graphed._graphs._seg) + len(graphed._graphs._enc)
                pg = graphed.step(a)
                n_graph = K_.LAUNCH_COUNT - n0 - n_eager
                if len(graphed._graphs._seg) + len(graphed._graphs._enc) == ncap:      # no capture in this step
                    assert n_graph == n_eager, 'graph replay must account for the same number of cutie_b200 kernels'
                assert float((eager.last_logits - graphed.last_logits).abs().max()) < 2e-4
            assert float((pe - pg).abs().max()) < 1e-4
            assert _sizes(eager) == _sizes(graphed)
    assert graphed._graphs is not None and len(graphed._graphs._seg) >= 1


def test_optimize_for_inference_keeps_parity():
    """BN folding + channels-last trunks + CUDA graphs vs the plain eager model (the fused conv epilogues the
    bench adds on top are compared with this configuration in tests/test_gpu_zz_conv_epilogues.py)."""
    from cutie_b200.config import default_config
    from cutie_b200.inference.inference_core import InferenceCore
    from oracle.synth import synthetic_video
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = default_config(mem_every=2, max_mem_frames=3)
    plain, fast = _net(cfg), _net(cfg).optimize_for_inference(fuse_epilogues=False, fuse_glue=False)
    a, b = InferenceCore(plain, cfg=cfg), InferenceCore(fast, cfg=cfg, use_cuda_graphs=True)
    frames, mask = synthetic_video(5, 96, 160, 3, seed=3)
    with torch.inference_mode():
        for ti in range(5):
            x = frames[ti].cuda()
            if ti == 0:
                a.step(x, mask.cuda(), objects=[1, 2, 3]); b.step(x, mask.cuda(), objects=[1, 2, 3])
            else:
                pa, pb = a.step(x), b.step(x)
                assert float((a.last_logits - b.last_logits).abs().max()) < 1e-3
                assert float((pa - pb).abs().max()) < 1e-3
