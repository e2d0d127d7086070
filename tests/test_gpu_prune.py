"""Arena pruning (device option prune_tokens = PruneActiveTokens every prune_interval frames,
online-decoder-base-inl.h:438-480 called at :660-661): results must not change — one-best and raw
lattice bit-identical to the unpruned decoder and to the canonical oracle — while the token arena
stays bounded by the live tokens instead of growing with the utterance."""
import ctypes as C

import numpy as np
import pytest

from asr_decoder_b200 import _lib, synth
from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, LatticeFasterDecoderConfig

pytestmark = pytest.mark.gpu


def _counters(dec):
    L = _lib.lib()
    ae, aa, tk = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    _lib.check(L.asrd_get_counters(dec.handles, dec.n, C.byref(ae), C.byref(aa), C.byref(tk), None), "counters")
    return dict(arcs=ae.value, tokens=tk.value, pruned=int(L.asrd_last_pruned_tokens()),
                peak=int(L.asrd_last_peak_tokens()))


def _same_lattice(a, b):
    ta, la = a
    tb, lb = b
    assert len(ta) == len(tb) and len(la) == len(lb)
    assert ta.tobytes() == tb.tobytes()
    assert la.tobytes() == lb.tobytes()


def _chunks(ll, step):
    return [ll[i:i + step] for i in range(0, ll.shape[0], step)]


KEPT = {}


# k_prune (pull sweep, map in shared memory) swept to frame 0 every time / k_lattice PRUNE mode (HBM
# maps, sweeps until nothing changes) / k_prune at its default depth (prune_interval frames below
# the previous frontier)
@pytest.mark.parametrize("kernel,depth", [("1", "-1"), ("0", "-1"), ("1", None)])
@pytest.mark.parametrize("chunk,interval", [(1000, 25), (30, 25), (7, 10)])
def test_pruned_decoder_gives_identical_results(monkeypatch, chunk, interval, kernel, depth):
    monkeypatch.setenv("ASRD_PRUNE_KERNEL", kernel)
    if depth is not None:
        monkeypatch.setenv("ASRD_PRUNE_DEPTH", depth)
    fst = synth.make_graph(20000, 5.0, 200, seed=77)
    T = 160
    lls = [synth.make_loglikes(T, 200, 2.0 if i % 2 == 0 else 1.3, seed=600 + i) for i in range(6)]
    cfg = LatticeFasterDecoderConfig(beam=13.0, max_active=3000, min_active=100, lattice_beam=6.0,
                                     prune_interval=interval)
    g = CudaFst(fst)
    plain = CudaDecoderBatch(g, cfg, len(lls), max_frames=T + 8, collect_stats=True)
    want = plain.Decode(lls)
    want_lat = [plain.GetRawLattice(i) for i in range(len(lls))]
    c0 = _counters(plain)
    assert c0["pruned"] == 0

    pruned = CudaDecoderBatch(g, cfg, len(lls), max_frames=T + 8, collect_stats=True, prune_tokens=True)
    for rep in range(2):  # the second utterance on the same objects starts from a clean slate
        pruned.InitDecoding()
        for k in range(0, T, chunk):
            pruned.AdvanceDecoding([ll[k:k + chunk] for ll in lls])
            if chunk == 30 and k == 60:  # a partial result in mid-utterance, between two prunes
                mid = pruned.GetBestPath(False)
                assert all(m.ok for m in mid)
        pruned.FinalizeDecoding()
        got = pruned.GetBestPath(True)
        for w, x in zip(want, got):
            assert x.ok == w.ok and x.words == w.words and x.ali == w.ali
            assert np.float32(x.tot).view(np.uint32) == np.float32(w.tot).view(np.uint32)
        c1 = _counters(pruned)   # (before GetRawLattice: it brings the last frames down to their survivors first)
        for i in range(len(lls)):
            _same_lattice(want_lat[i], pruned.GetRawLattice(i))
            # per-frame statistics are taken when the frame is decoded: the pruning cannot touch them
            a, b = plain.frame_stats(i), pruned.frame_stats(i)
            for k in a.dtype.names:
                if k != "arcs_admitted":   # (counted against the RUNNING cutoff: depends on the warp schedule)
                    assert a[k].tobytes() == b[k].tobytes(), k
        assert c1["arcs"] == c0["arcs"]
        assert c1["pruned"] > 0.5 * c0["tokens"], (c1, c0)      # most tokens never reach the lattice
        assert c1["tokens"] + c1["pruned"] == c0["tokens"]
        assert all(pruned.status(i) == 0 for i in range(len(lls)))
        if depth == "-1":  # swept all the way, the two prune kernels leave the same arena behind
            assert KEPT.setdefault((chunk, interval), c1["tokens"]) == c1["tokens"]


def test_arena_is_bounded_by_the_live_tokens():
    """With pruning the arena only has to hold the frames since the last prune plus the pruned
    history: an utterance decodes in an arena an unpruned decoder overflows."""
    fst = synth.make_graph(20000, 5.0, 200, seed=78)
    T = 400
    ll = synth.make_loglikes(T, 200, 2.0, seed=901)
    cfg = LatticeFasterDecoderConfig(beam=13.0, max_active=3000, min_active=100, lattice_beam=6.0, prune_interval=25)
    g = CudaFst(fst)
    big = CudaDecoderBatch(g, cfg, 1, max_frames=T + 8)
    want = big.Decode([ll])[0]
    total = _counters(big)["tokens"]
    cap = total // 4
    small = CudaDecoderBatch(g, cfg, 1, max_frames=T + 8, token_capacity=cap)
    assert not small.Decode([ll])[0].ok and small.status(0) != 0      # overflow is reported, never silent
    pr = CudaDecoderBatch(g, cfg, 1, max_frames=T + 8, token_capacity=cap, prune_tokens=True)
    pr.InitDecoding()
    for c in _chunks(ll, 25):
        pr.AdvanceDecoding([c])
    pr.FinalizeDecoding()
    got = pr.GetBestPath(True)[0]
    assert got.ok and got.words == want.words and got.ali == want.ali
    c = _counters(pr)
    assert c["peak"] <= cap and c["tokens"] < total // 8, (c, total)


def test_prune_options_are_validated():
    fst = synth.make_graph(500, 4.0, 40, seed=3)
    g = CudaFst(fst)
    L = _lib.lib()
    cfg = LatticeFasterDecoderConfig(beam=10.0, max_active=500, min_active=20, lattice_beam=5.0).to_c()
    cfg.prune_interval = 0
    opts = _lib.asrd_device_options(0, 0, 32, 0, 0, 1)
    h = C.c_void_p()
    assert L.asrd_decoder_create(g.h, C.byref(cfg), C.byref(opts), C.byref(h)) == -1  # ASRD_ERR_BAD_ARG


def test_hub_states_and_mixed_kernels(monkeypatch):
    """Destinations with thousands of incoming arcs (the super-final state: one eps arc from every
    final state; here also an emitting hub every 7th state points at) are walked by the whole CTA
    in k_prune; and with a small on-chip budget (ASRD_PRUNE_CAP) the streams whose frames do not fit
    go to the HBM-map sweep while the others stay on chip — the answers never change."""
    import numpy as np
    from asr_decoder_b200.fstio import Fst
    base = synth.make_graph(6000, 5.0, 120, seed=5, p_final=0.3)
    arcs = base.arcs.copy()
    # redirect the first emitting non-loop arc of every 7th state to one hub state
    off = np.concatenate([[0], np.cumsum(base.num_arcs.astype(np.int64))])
    hub = 4321
    for s in range(0, 6000, 7):
        for a in range(off[s] + int(base.niepsilons[s]) + 1, off[s + 1]):
            arcs["nextstate"][a] = hub
            break
    fst = Fst(start=base.start, final_state=base.final_state, arcs=arcs, num_arcs=base.num_arcs,
              niepsilons=base.niepsilons, noepsilons=base.noepsilons)
    T = 120
    lls = [synth.make_loglikes(T, 120, 1.5 + 0.5 * (i % 2), seed=40 + i) for i in range(4)]
    cfg = LatticeFasterDecoderConfig(beam=14.0, max_active=4000, min_active=100, lattice_beam=7.0, prune_interval=20)
    g = CudaFst(fst)
    plain = CudaDecoderBatch(g, cfg, len(lls), max_frames=T + 8)
    want = plain.Decode(lls)
    want_lat = [plain.GetRawLattice(i) for i in range(len(lls))]
    for cap in (None, "512"):
        if cap:
            monkeypatch.setenv("ASRD_PRUNE_CAP", cap)
        pr = CudaDecoderBatch(g, cfg, len(lls), max_frames=T + 8, prune_tokens=True)
        pr.InitDecoding()
        for k in range(0, T, 20):
            pr.AdvanceDecoding([ll[k:k + 20] for ll in lls])
        pr.FinalizeDecoding()
        got = pr.GetBestPath(True)
        for w, x in zip(want, got):
            assert x.ok == w.ok and x.words == w.words and x.ali == w.ali and x.tot_bits == w.tot_bits
        for i in range(len(lls)):
            _same_lattice(want_lat[i], pr.GetRawLattice(i))
        assert _counters(pr)["pruned"] > 0


def test_long_utterance_in_a_bounded_arena():
    """1500 frames (45 s of audio) through an arena sized for ~60 frames: what pruning is for."""
    fst = synth.make_graph(50000, 5.0, 300, seed=8)
    T = 1500
    ll = synth.make_loglikes(T, 300, 2.0, seed=77)
    cfg = LatticeFasterDecoderConfig(beam=13.0, max_active=3000, min_active=100, lattice_beam=6.0, prune_interval=25)
    g = CudaFst(fst)
    big = CudaDecoderBatch(g, cfg, 1, max_frames=T + 8, token_capacity=T * 6000)
    want = big.Decode([ll])[0]
    total = _counters(big)["tokens"]
    cap = 60 * 5000
    assert total > 8 * cap
    pr = CudaDecoderBatch(g, cfg, 1, max_frames=T + 8, token_capacity=cap, prune_tokens=True)
    pr.InitDecoding()
    for c in _chunks(ll, 30):
        pr.AdvanceDecoding([c])
    pr.FinalizeDecoding()
    got = pr.GetBestPath(True)[0]
    assert got.ok and got.words == want.words and got.ali == want.ali and got.tot_bits == want.tot_bits
    c = _counters(pr)
    assert c["peak"] <= cap, c
    _same_lattice(big.GetRawLattice(0), pr.GetRawLattice(0))
