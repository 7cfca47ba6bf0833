"""The CUDA-graph frame path and the encoder look-ahead, driven on the CPU with stand-in graphs.

A stand-in "captured graph" keeps the static-buffer semantics that matter for correctness: inputs are copied into fixed
tensors, a replay recomputes and OVERWRITES the fixed output tensors in place (same storage, same data_ptr -- the
segment / mask-encoder graphs are keyed on those pointers).  Stream operations are no-ops, so look-ahead work executes
at the moment it is enqueued -- the earliest legal schedule: the next frame's encoder runs BEFORE this frame's memory
read and decoder.  If the look-ahead ever wrote into buffers the current frame still reads, or a step consumed features
of the wrong frame, the logits would differ from the eager model."""
import contextlib

import pytest
import torch


def _copy_tree_(dst, src):
    if isinstance(dst, torch.Tensor):
        dst.copy_(src)
    elif dst is not None:
        for d, s in zip(dst, src):
            _copy_tree_(d, s)


class _FakeCaptured:
    replays = 0

    def __init__(self, fn, static_inputs):
        self.fn, self.inputs = fn, static_inputs
        self.outputs = fn(*static_inputs)
        self.kernel_launches = 0

    def replay(self):
        _FakeCaptured.replays += 1
        _copy_tree_(self.outputs, self.fn(*self.inputs))
        return self.outputs


class _NoStreams:
    def side_wait_main(self, device): pass
    def keep_alive_on_side(self, tensor): pass
    def on_side(self, device): return contextlib.nullcontext()
    def record_on_side(self, device): return object()
    def main_wait_event(self, ev): pass


@pytest.fixture
def graph_path_on_cpu(monkeypatch, cpu_kernels):
    import cutie_b200.inference.frame_graphs as fg
    import cutie_b200.inference.inference_core as ic
    monkeypatch.setattr(fg, '_Captured', _FakeCaptured)
    monkeypatch.setattr(ic, '_graphable', lambda t: True)
    monkeypatch.setattr(ic, '_CudaStreamOps', _NoStreams)
    _FakeCaptured.replays = 0
    return ic


def _net(cfg, optimise):
    from cutie_b200.model.cutie import CUTIE
    from oracle.synth import synthetic_state_dict
    n = CUTIE(cfg).eval()
    n.load_state_dict(synthetic_state_dict(n.state_dict(), 0))
    return n.optimize_for_inference() if optimise else n


@pytest.mark.parametrize('lookahead', [False, True])
def test_graph_path_with_and_without_lookahead_matches_eager(graph_path_on_cpu, lookahead):
    ic = graph_path_on_cpu
    from cutie_b200.config import default_config
    from oracle.synth import synthetic_video
    cfg = default_config(mem_every=3, max_mem_frames=3)
    net = _net(cfg, False)
    eager = ic.InferenceCore(net, cfg=cfg)
    graphed = ic.InferenceCore(net, cfg=cfg, use_cuda_graphs=True)
    T = 11
    frames, mask = synthetic_video(T + 1, 96, 160, 3, seed=4)
    hits = 0
    with torch.inference_mode():
        for ti in range(T):
            kw = dict(objects=[1, 2, 3]) if ti == 0 else {}
            args = (frames[ti], mask) if ti == 0 else (frames[ti],)
            # every fourth announcement names a different tensor: the look-ahead must be discarded
            nxt = None if not lookahead else (frames[ti + 1] if ti % 4 != 3 else frames[ti + 1].clone())
            pe = eager.step(*args, **kw)
            pg = graphed.step(*args, next_image=nxt, **kw)
            assert torch.allclose(pg, pe, atol=1e-6), (ti, float((pg - pe).abs().max()))
            if ti > 0:
                assert torch.allclose(graphed.last_logits, eager.last_logits, atol=1e-5)
            assert graphed.memory.work_mem.size(0) == eager.memory.work_mem.size(0)
    assert _FakeCaptured.replays > T                                   # encoder + segment (+ mask encoder) replays
    g = graphed._graphs
    assert len(g._enc) == (2 if lookahead else 1)                      # the second capture slot exists only with look-ahead
    assert len(g._seg) >= (2 if lookahead else 1)                      # keyed on the slot's buffers


def test_lookahead_outputs_are_not_overwritten_while_in_use(graph_path_on_cpu):
    """Direct check of the slot discipline: the features a step works on keep their values until the step returns, even
    though the next frame's encoder has already run (stand-in streams execute look-ahead work immediately)."""
    ic = graph_path_on_cpu
    from cutie_b200.config import default_config
    from oracle.synth import synthetic_video
    cfg = default_config(mem_every=2, max_mem_frames=3)
    net = _net(cfg, False)
    proc = ic.InferenceCore(net, cfg=cfg, use_cuda_graphs=True)
    frames, mask = synthetic_video(6, 96, 160, 2, seed=6)
    seen = {}
    orig = proc._segment

    def spy(key, selection, pix_feat, ms_features, update_sensory=True):
        with torch.inference_mode():
            ref_ms, ref_pix = net.encode_image(seen['image'])
        assert torch.allclose(pix_feat, ref_pix, atol=1e-6), 'segment() was handed the features of another frame'
        return orig(key, selection, pix_feat, ms_features, update_sensory=update_sensory)
    proc._segment = spy
    from cutie_b200.utils.tensor_utils import pad_divide_by
    with torch.inference_mode():
        for ti in range(5):
            seen['image'] = pad_divide_by(frames[ti], 16)[0].unsqueeze(0)
            if ti == 0:
                proc.step(frames[0], mask, objects=[1, 2], next_image=frames[1])
            else:
                proc.step(frames[ti], next_image=frames[ti + 1])
