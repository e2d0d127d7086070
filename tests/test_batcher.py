"""Streaming batcher (asr_decoder_b200/batcher.py).  CPU: the scheduling
logic against a stub engine that records the batched calls — channel reuse, max-batch-size,
oldest-first, one chunk per stream per step, finalisation exactly once.  GPU: streams that arrive
and end at different times decode exactly like the offline batch."""
import numpy as np
import pytest

from asr_decoder_b200.batcher import StreamingBatcher


class StubEngine:
    def __init__(self):
        self.log = []
        self.frames = {}

    def init(self, ch):
        self.log.append(("init", tuple(ch)))
        for c in ch:
            self.frames[c] = 0

    def advance(self, ch, chunks):
        assert len(set(ch)) == len(ch)
        self.log.append(("advance", tuple(ch), tuple(x.shape[0] for x in chunks)))
        for c, x in zip(ch, chunks):
            self.frames[c] += x.shape[0]

    def finalize(self, ch):
        self.log.append(("finalize", tuple(ch)))

    def best_paths(self, ch):
        return [("path", c, self.frames[c]) for c in ch]


def rows(t):
    return np.zeros((t, 4), np.float32)


def test_scheduling_rules():
    eng = StubEngine()
    b = StreamingBatcher(eng, num_channels=3, max_batch_size=2, frames_per_chunk=30)
    assert b.open("a") and b.open("b") and b.open("c")
    assert not b.open("d")                       # every channel is busy
    b.push("a", rows(70))                        # cut into 30 + 30 + 10
    b.push("b", rows(30), last=True)
    b.push("c", rows(45))
    out = b.step()                               # init a, b, c; advance the two oldest: a and b; b is complete
    assert eng.log[0] == ("init", (0, 1, 2))
    assert eng.log[1] == ("advance", (0, 1), (30, 30))
    assert eng.log[2] == ("finalize", (1,))
    assert [s for s, _ in out] == ["b"] and out[0][1] == ("path", 1, 30)
    assert b.step() == []
    assert eng.log[3] == ("advance", (2, 0), (30, 30))          # c waited longest, then a
    assert b.open("d")                           # b's channel is free again
    b.push("d", rows(5), last=True)
    b.push("a", rows(1), last=True)
    b.push("c", rows(0), last=True)              # an empty last chunk just ends the stream
    out = b.drain()
    assert sorted(s for s, _ in out) == ["a", "c", "d"]
    frames = {s: p[2] for s, p in out}
    assert frames == {"a": 71, "c": 45, "d": 5}
    assert not b.busy()
    assert all(len(e[1]) <= 2 for e in eng.log if e[0] == "advance")
    assert sum(1 for e in eng.log if e[0] == "finalize" for _ in e[1]) == 4
    with pytest.raises(KeyError):
        b.open("x") and b.open("x")


@pytest.mark.gpu
def test_ragged_arrivals_equal_offline_decode():
    from asr_decoder_b200 import synth
    from asr_decoder_b200.batcher import make_cuda_batcher
    from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, LatticeFasterDecoderConfig
    fst = synth.make_graph(30000, 5.0, 300, seed=12)
    g = CudaFst(fst)
    cfg = LatticeFasterDecoderConfig(beam=13.0, max_active=3000, min_active=200, lattice_beam=8.0)
    rng = np.random.default_rng(3)
    utts = [synth.make_loglikes(int(rng.integers(20, 140)), 300, 2.0 + 0.5 * (i % 3), seed=300 + i) for i in range(40)]
    offline = CudaDecoderBatch(g, cfg, len(utts), max_frames=160).Decode(utts)
    b = make_cuda_batcher(g, cfg, num_channels=16, max_batch_size=12, frames_per_chunk=30, max_frames=160)
    pending, fed, results = list(range(len(utts))), {}, {}
    while pending or b.busy():
        while pending and b.open(pending[0]):    # new streams take free channels
            fed[pending.pop(0)] = 0
        for s in list(fed):                      # every open stream delivers its next (ragged) packet
            if fed[s] is None:
                continue
            n = int(rng.integers(10, 70))
            f0, f1 = fed[s], min(fed[s] + n, utts[s].shape[0])
            b.push(s, utts[s][f0:f1], last=f1 == utts[s].shape[0])
            fed[s] = None if f1 == utts[s].shape[0] else f1
        for s, path in b.step():
            results[s] = path
    assert sorted(results) == list(range(len(utts)))
    for i, ref in enumerate(offline):
        got = results[i]
        assert got.ok and ref.ok and got.words == ref.words and got.ali == ref.ali and got.tot_bits == ref.tot_bits, i
    assert max(b.stats["batch_sizes"]) <= 12
