"""Happens-before model check of the encoder look-ahead protocol (cutie_b200/inference/inference_core.EncoderLookahead).

The real thing needs two CUDA streams; here the stream operations are replaced by a recorder that builds the
happens-before graph (program order per stream + wait_stream / wait_event edges), the encoder "graph replay" is a node
that WRITES a capture slot, and every consumer of a frame's features is a node that READS it.  For random schedules of
announced / mis-announced / unannounced frames the test asserts that (1) a step always gets the features of ITS frame,
(2) every read of a slot is ordered after the write it expects and (3) no write to a slot is concurrent with -- or
wedged between -- a read of that slot and the write the read expects."""
import contextlib
import random

import pytest
import torch

from cutie_b200.inference.inference_core import EncoderLookahead


class Recorder:
    def __init__(self):
        self.nodes = []                  # (stream, kind, slot, frame)
        self.edges = set()               # (a, b): a happens-before b
        self.last = {'main': None, 'side': None}
        self.pending_dep = {'main': [], 'side': []}
        self.cur = 'main'

    def node(self, kind, slot=None, frame=None):
        i = len(self.nodes)
        self.nodes.append((self.cur, kind, slot, frame))
        if self.last[self.cur] is not None:
            self.edges.add((self.last[self.cur], i))
        for d in self.pending_dep[self.cur]:
            self.edges.add((d, i))
        self.pending_dep[self.cur] = []
        self.last[self.cur] = i
        return i

    # -- the _CudaStreamOps interface --
    def side_wait_main(self, device):
        if self.last['main'] is not None:
            self.pending_dep['side'].append(self.last['main'])

    def keep_alive_on_side(self, tensor):
        pass

    @contextlib.contextmanager
    def on_side(self, device):
        prev, self.cur = self.cur, 'side'
        try:
            yield
        finally:
            self.cur = prev

    def record_on_side(self, device):
        return self.last['side']

    def main_wait_event(self, ev):
        if ev is not None:
            self.pending_dep['main'].append(ev)

    # -- reachability --
    def hb(self):
        n = len(self.nodes)
        reach = [set() for _ in range(n)]
        succ = [[] for _ in range(n)]
        for a, b in self.edges:
            succ[a].append(b)
        for i in range(n - 1, -1, -1):           # edges always go forward in recording order
            for j in succ[i]:
                reach[i].add(j)
                reach[i] |= reach[j]
        return reach


def _run(seed: int, steps: int = 40):
    rng = random.Random(seed)
    rec = Recorder()
    frames = [torch.zeros(3, 4, 4) + i for i in range(steps + 2)]

    def encode(image, slot):
        f = int(image.flatten()[0])
        rec.node('write', slot, f)
        return {'slot': slot, 'frame': f}
    la = EncoderLookahead(encode, ops=rec)
    prep = lambda t: t.unsqueeze(0)
    for t in range(steps):
        img = frames[t].unsqueeze(0)
        out, hit = la.current(t, img, frames[t])
        assert out['frame'] == t, f'step {t} got the features of frame {out["frame"]}'
        mode = rng.choice(['announce', 'announce', 'announce', 'none', 'wrong_tensor', 'wrong_frame'])
        if mode == 'announce':
            la.ahead(t, frames[t + 1], prep)
        elif mode == 'wrong_tensor':             # same content, another tensor: must be discarded
            la.ahead(t, frames[t + 1].clone(), prep)
        elif mode == 'wrong_frame':              # announces a frame that will not come next
            la.ahead(t, frames[t + 2] if t + 2 < len(frames) else frames[0], prep)
        for _ in range(rng.randint(1, 3)):       # memory read / segment graph / mask-encoder graph of this frame
            rec.node('read', out['slot'], t)
    return rec


@pytest.mark.parametrize('seed', range(12))
def test_lookahead_protocol_has_no_races(seed):
    rec = _run(seed)
    reach = rec.hb()
    writes = [(i, n[2], n[3]) for i, n in enumerate(rec.nodes) if n[1] == 'write']
    reads = [(i, n[2], n[3]) for i, n in enumerate(rec.nodes) if n[1] == 'read']
    assert any(rec.nodes[i][0] == 'side' for i, _, _ in writes), 'schedule never used the side stream'
    for r, slot, frame in reads:
        mine = [w for w, s, f in writes if s == slot and f == frame and r in reach[w]]
        assert mine, f'read {r} (slot {slot}, frame {frame}) is not ordered after a write of its frame'
        w0 = max(mine)
        for w, s, f in writes:
            if s != slot or w == w0:
                continue
            before, after = r in reach[w], w in reach[r]
            assert before or after, f'write {w} (frame {f}) is concurrent with read {r} of slot {slot}'
            if before and f != frame:
                assert w0 in reach[w], f'write {w} (frame {f}) can land between write {w0} and read {r}'


def test_hit_and_miss_bookkeeping():
    rec = Recorder()
    log = []

    def encode(image, slot):
        log.append((rec.cur, slot))
        return ('features', slot)
    la = EncoderLookahead(encode, ops=rec)
    a, b, c = torch.zeros(3, 2, 2), torch.ones(3, 2, 2), torch.ones(3, 2, 2) * 2
    sid = lambda t: t                    # a hit is decided on the announced tensor's memory and version counter
    out, hit = la.current(0, a.unsqueeze(0), sid(a))
    assert not hit and log == [('main', 0)]
    la.ahead(0, b, lambda t: t.unsqueeze(0))
    assert log[-1] == ('side', 1)
    out, hit = la.current(1, b.unsqueeze(0), sid(b))
    assert hit and out == ('features', 1) and la.slot == 1 and len(log) == 2          # no re-encode
    la.ahead(1, c, lambda t: t.unsqueeze(0))
    assert log[-1] == ('side', 0)
    out, hit = la.current(2, b.unsqueeze(0), sid(b))                                   # c was announced, b arrives
    assert not hit and log[-1] == ('main', 1) and la.slot == 1                          # re-encoded into the current slot
    out, hit = la.current(3, c.unsqueeze(0), sid(c))                                   # nothing pending
    assert not hit and log[-1] == ('main', 1)
    # the announced buffer is overwritten in place before the next step (a reused staging buffer): never a hit
    la.ahead(3, a, lambda t: t.unsqueeze(0))
    a.add_(1.0)
    out, hit = la.current(4, a.unsqueeze(0), sid(a))
    assert not hit and log[-1][0] == 'main'
    # a different tensor of the same shape is never a hit either (the announced one is kept alive: no recycled address)
    la.ahead(4, b, lambda t: t.unsqueeze(0))
    out, hit = la.current(5, b.clone().unsqueeze(0), sid(b.clone()))
    assert not hit
