"""Pins the CPU oracle (oracle/) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  If this fails the oracle is wrong, not the kernels."""
import os

import numpy as np
import pytest
import torch

from oracle import memory_math as mm
from oracle import transformer as otf
from tests.conftest import GOLDEN


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope='module')
def kat():
    return np.load(os.path.join(GOLDEN, 'kat_memory.npz'))


def test_similarity_matches_reference(kat):
    mk, ms, qk, qe = (T(kat[k]) for k in ('mk', 'ms', 'qk', 'qe'))
    assert torch.allclose(mm.similarity_expanded(mk, ms, qk, qe), T(kat['sim']), rtol=1e-5, atol=1e-4)
    assert torch.allclose(mm.similarity_expanded(mk, ms, qk, None), T(kat['sim_no_qe']), rtol=1e-5, atol=1e-4)
    assert torch.allclose(mm.similarity_expanded(mk, None, qk, qe), T(kat['sim_no_ms']), rtol=1e-5, atol=1e-4)
    # the cancellation-free float64 form agrees with the reference's fp32 expansion to its rounding noise
    d = mm.similarity_direct(mk, ms, qk, qe)
    assert float((d - T(kat['sim']).double()).abs().max()) < 2e-4


def test_topk_softmax_usage_readout_match_reference(kat):
    sim, v = T(kat['sim']), T(kat['v'])
    idx, w = mm.topk_softmax(sim, 30)
    assert (idx.sort(1)[0] == T(kat['topk_idx']).sort(1)[0]).all()
    aff = mm.scatter_affinity(idx, w, sim.shape[1])
    assert torch.allclose(aff, T(kat['affinity']), atol=1e-7)
    assert (aff > 0).sum(1).eq(30).all()
    assert torch.allclose(mm.usage_from_affinity(aff), T(kat['usage']), atol=1e-6)
    assert torch.allclose(mm.readout(aff, v), T(kat['readout']), rtol=1e-5, atol=1e-5)
    assert torch.allclose(mm.sparse_readout(idx, w, v), T(kat['readout']), rtol=1e-4, atol=2e-5)
    assert torch.allclose(mm.dense_softmax(sim), T(kat['dense_affinity']), atol=1e-7)


def test_consolidation_matches_reference(kat):
    pk, pv, ps, _ = mm.consolidate(T(kat['cons_key']), T(kat['cons_shrinkage']), T(kat['cons_selection']),
                                   {1: T(kat['cons_v1']), 5: T(kat['cons_v5'])}, T(kat['cons_usage']), 16)
    assert torch.equal(pk, T(kat['cons_pk']))
    assert torch.allclose(pv[1], T(kat['cons_pv1']), rtol=1e-5, atol=1e-5)
    assert torch.allclose(pv[5], T(kat['cons_pv5']), rtol=1e-5, atol=1e-5)
    assert torch.allclose(ps, T(kat['cons_ps']), rtol=1e-5, atol=1e-5)


def test_tie_case_values(kat):
    aff = mm.scatter_affinity(*mm.topk_softmax(T(kat['tie_sim']), 30), T(kat['tie_sim']).shape[1])
    ref = T(kat['tie_affinity'])
    assert torch.allclose(aff.sort(1, descending=True)[0][:, :30], ref.sort(1, descending=True)[0][:, :30], atol=1e-7)


def _weights():
    from cutie_b200.config import default_config
    from cutie_b200.model.cutie import CUTIE
    from oracle.synth import synthetic_state_dict
    cfg = default_config()
    net = CUTIE(cfg).eval()
    net.load_state_dict(synthetic_state_dict(net.state_dict(), 0))
    return cfg, net


def test_query_transformer_matches_reference():
    g = np.load(os.path.join(GOLDEN, 'qt_module.npz'))
    _, net = _weights()
    sd = net.state_dict()
    trace = {}
    with torch.inference_mode():
        out, logits = otf.query_transformer(T(g['pixel']), T(g['obj_summaries']), sd, trace=trace)
    assert torch.allclose(otf.sinusoid_pe(6, 10), T(g['pe']), atol=1e-6)
    assert torch.allclose(out, T(g['out']), rtol=1e-4, atol=2e-4)
    for i in range(4):
        assert torch.allclose(logits[i], T(g[f'aux_logits_{i}']), rtol=1e-4, atol=2e-4)
    for i in range(3):
        for name in ('after_rfp', 'query', 'pixel_flat', 'pixel'):
            ref = T(g[f'b{i}_{name}'])
            got = trace[f'b{i}_{name}'].reshape(ref.shape)
            assert torch.allclose(got, ref, rtol=1e-4, atol=3e-4), (i, name)
    # attention mask semantics (object_transformer.py:179-205)
    blocked = otf.attention_block_mask(otf.foreground_map(T(g['aux_logits_0'])), 16)
    ref_mask = T(g['b0_attn_mask_in'])                      # [(B*K*heads), Q, HW]
    assert torch.equal(blocked.repeat_interleave(8, 0), ref_mask)


LT_SMALL = dict(max_mem_frames=4, min_mem_frames=2, num_prototypes=16, max_num_tokens=60, buffer_tokens=20)


@pytest.mark.parametrize('name,over,T_,K', [
    ('fifo', dict(mem_every=2, max_mem_frames=3), 10, 3),
    ('longterm', dict(mem_every=1, use_long_term=True, long_term=LT_SMALL), 14, 2),
])
def test_full_frame_oracle_matches_reference(name, over, T_, K):
    from cutie_b200.config import default_config
    from cutie_b200.model.cutie import CUTIE
    from oracle.cpu_core import OracleCore
    from oracle.synth import synthetic_state_dict, synthetic_video
    g = np.load(os.path.join(GOLDEN, f'e2e_{name}.npz'))
    cfg = default_config(**over)
    net = CUTIE(cfg).eval()
    net.load_state_dict(synthetic_state_dict(net.state_dict(), 0))
    frames, mask = synthetic_video(T_, 96, 160, K, seed=3)
    oc = OracleCore(net, cfg)
    li = 0
    with torch.inference_mode():
        for ti in range(T_):
            prob = oc.step(frames[ti], mask, objects=list(range(1, K + 1))) if ti == 0 else oc.step(frames[ti])
            row = []
            for b in sorted(oc.work.buckets):
                row += [b, oc.work.size(b), oc.work.perm_end.get(b, 0), oc.long.size(b) if oc.use_long_term else 0]
            assert row == [int(x) for x in g['sizes'][ti] if x >= 0]
            if ti > 0:
                assert float(np.abs(oc.last_logits.numpy() - g['logits'][li:li + 1]).max()) < 2e-4
                li += 1
            assert (oc.output_prob_to_mask(prob).numpy() == g['masks'][ti]).mean() > 0.999
    assert float((prob - T(g['final_prob'])).abs().max()) < 1e-4


def test_oracle_buckets_and_delete():
    from cutie_b200.config import default_config
    from cutie_b200.model.cutie import CUTIE
    from oracle.cpu_core import OracleCore
    from oracle.synth import synthetic_state_dict, synthetic_video
    g = np.load(os.path.join(GOLDEN, 'e2e_buckets.npz'))
    cfg = default_config(mem_every=2, max_mem_frames=3)
    net = CUTIE(cfg).eval()
    net.load_state_dict(synthetic_state_dict(net.state_dict(), 0))
    frames, _ = synthetic_video(8, 96, 160, 3, seed=3)
    oc = OracleCore(net, cfg)
    with torch.inference_mode():
        for ti in range(8):
            if ti == 0:
                prob = oc.step(frames[0], T(g['first_mask']), objects=[1, 2])
            elif ti == 3:
                prob = oc.step(frames[3], T(g['second_mask']), objects=[7])
            elif ti == 6:
                oc.delete_objects([1])
                prob = oc.step(frames[6])
            else:
                prob = oc.step(frames[ti])
            if ti == 4:
                assert float(np.abs(oc.last_logits.numpy() - g['logits_f4']).max()) < 2e-4
            assert (oc.output_prob_to_mask(prob).numpy() == g['masks'][ti]).mean() > 0.999
    assert float(np.abs(oc.last_logits.numpy() - g['logits']).max()) < 2e-4
