"""GPU parity: the CUDA path (through the C ABI) against the canonical-mode oracle.

Bit-exact bar: one-best words, alignment and cost bits, per-frame token counts, cutoffs,
adaptive beams and emitting-arc counts.  The oracle is the checker only.
"""
import numpy as np
import pytest

from asr_decoder_b200 import synth
from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, LatticeFasterDecoderConfig

pytestmark = pytest.mark.gpu


def _cfg(**kw):
    base = dict(beam=13.0, max_active=7000, min_active=200, lattice_beam=8.0)
    base.update(kw)
    return LatticeFasterDecoderConfig(**base)


def _oracle_decode(O, og, cfg, ll, finalize=True):
    d = O.OracleDecoder(og, O.make_config(cfg.beam, cfg.max_active, cfg.min_active, cfg.lattice_beam,
                                          cfg.prune_interval, cfg.beam_delta, cfg.hash_ratio,
                                          cfg.prune_scale), O.MODE_CANONICAL)
    r = d.decode(ll, finalize=finalize)
    return r, d.frame_stats()


def _compare(bp, st, ref, rst, tag=""):
    assert bp.ok == ref.ok, tag
    assert bp.words == ref.words, tag
    assert bp.ali == ref.ali, tag
    assert bp.tot_bits == ref.tot_bits, (tag, bp.tot, ref.tot)
    assert np.array_equal(bp.ilabel, ref.ilabel) and np.array_equal(bp.olabel, ref.olabel), tag
    assert np.array_equal(bp.graph.view(np.uint32), ref.graph.view(np.uint32)), tag
    assert np.array_equal(bp.acoustic.view(np.uint32), ref.acoustic.view(np.uint32)), tag
    if st is not None:
        assert len(st) == len(rst), tag
        assert np.array_equal(st["n_tokens"], rst["n_raw"]), tag
        assert np.array_equal(rst["n_raw"], rst["n_within"]), tag
        for a, b in (("cur_cutoff", "cur_cutoff"), ("abeam", "abeam"), ("next_cutoff", "next_cutoff"),
                     ("best", "best")):
            assert np.array_equal(st[a].view(np.uint32), rst[b].view(np.uint32)), (tag, a)
        assert np.array_equal(st["n_in"], rst["n_in"]), tag
        assert np.array_equal(st["arcs_expanded"].astype(np.int64), rst["arcs_expanded"]), tag


@pytest.mark.parametrize("sigma", [2.0, 3.0, 1.5])
def test_config1_single_stream(oracle_mod, sigma):
    """BASELINE.json configs[0]: 10k states / 50k arcs / 200 pdfs, 500 frames, beam 13, max-active 7000."""
    O = oracle_mod
    fst = synth.make_graph(10000, 5.0, 200, seed=12345)
    ll = synth.make_loglikes(500, 200, sigma, seed=777)
    cfg = _cfg()
    g = CudaFst(fst)
    dec = CudaDecoderBatch(g, cfg, 1, max_frames=512, collect_stats=True)
    bp = dec.Decode([ll])[0]
    assert dec.status(0) == 0
    ref, rst = _oracle_decode(O, O.OracleGraph(fst), cfg, ll)
    _compare(bp, dec.frame_stats(0), ref, rst, f"sigma={sigma}")


def test_batch_of_streams_ragged(oracle_mod):
    """Several streams of different lengths in one batch, each equal to its own oracle decode."""
    O = oracle_mod
    fst = synth.make_graph(5000, 5.0, 120, seed=99)
    cfg = _cfg(max_active=3000)
    og = O.OracleGraph(fst)
    lens = [37, 120, 1, 64, 200, 90, 5, 150]
    lls = [synth.make_loglikes(t, 120, 2.0 + 0.25 * (i % 3), seed=1000 + i) for i, t in enumerate(lens)]
    g = CudaFst(fst)
    dec = CudaDecoderBatch(g, cfg, len(lens), max_frames=256, collect_stats=True)
    out = dec.Decode(lls)
    for i, bp in enumerate(out):
        assert dec.status(i) == 0
        ref, rst = _oracle_decode(O, og, cfg, lls[i])
        _compare(bp, dec.frame_stats(i), ref, rst, f"stream {i}")


def test_chunked_advance_equals_one_shot(oracle_mod):
    """Streaming: 30-frame chunks (config 5's call pattern) give the same result as one call,
    and the decoder object is reusable across utterances (InitDecoding again)."""
    O = oracle_mod
    fst = synth.make_graph(8000, 5.0, 150, seed=5)
    cfg = _cfg(max_active=4000)
    ll = synth.make_loglikes(157, 150, 2.0, seed=31)
    g = CudaFst(fst)
    dec = CudaDecoderBatch(g, cfg, 1, max_frames=256, collect_stats=True)
    ref, rst = _oracle_decode(O, O.OracleGraph(fst), cfg, ll)
    for rep in range(2):
        dec.InitDecoding()
        for f0 in range(0, ll.shape[0], 30):
            dec.AdvanceDecoding([ll[f0:f0 + 30]])
        assert dec.NumFramesDecoded(0) == ll.shape[0]
        dec.FinalizeDecoding()
        bp = dec.GetBestPath()[0]
        _compare(bp, dec.frame_stats(0), ref, rst, f"rep {rep}")


@pytest.mark.parametrize("kw", [dict(max_active=500, min_active=50), dict(beam=6.0, min_active=2000, max_active=6000),
                                dict(beam=20.0, max_active=2 ** 31 - 1, min_active=0)])
def test_cutoff_branches(oracle_mod, kw):
    """max-active binding, min-active widening and beam-only GetCutoff branches (inl.h:188-232)."""
    O = oracle_mod
    fst = synth.make_graph(6000, 5.0, 100, seed=21)
    cfg = _cfg(**kw)
    ll = synth.make_loglikes(120, 100, 2.0, seed=8)
    g = CudaFst(fst)
    dec = CudaDecoderBatch(g, cfg, 1, max_frames=128, collect_stats=True, hash_capacity=1 << 16)
    bp = dec.Decode([ll])[0]
    assert dec.status(0) == 0
    oc = O.make_config(cfg.beam, min(cfg.max_active, 2 ** 31 - 1), cfg.min_active, cfg.lattice_beam)
    d = O.OracleDecoder(O.OracleGraph(fst), oc, O.MODE_CANONICAL)
    ref = d.decode(ll)
    _compare(bp, dec.frame_stats(0), ref, d.frame_stats(), str(kw))


@pytest.mark.parametrize("name", ["g1", "g2", "g3"])
def test_golden_fixtures_from_the_compiled_reference(oracle_mod, name):
    """CUDA path == the compiled reference's own one-best (committed golden vectors, all
    self-stable) and == the canonical oracle per frame."""
    import json
    import os
    from asr_decoder_b200 import fstio
    O = oracle_mod
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    fst = fstio.read_fst(os.path.join(gold, name + ".fst"))
    lls = fstio.read_loglikes(os.path.join(gold, name + ".llb"))
    meta = json.load(open(os.path.join(gold, name + ".json")))
    cfg = _cfg(**meta["config"])
    g = CudaFst(fst)
    dec = CudaDecoderBatch(g, cfg, len(lls), max_frames=128, collect_stats=True)
    out = dec.Decode(lls)
    og = O.OracleGraph(fst)
    for i, (bp, ref) in enumerate(zip(out, meta["reference"])):
        assert dec.status(i) == 0
        assert meta["self_stable"][i]
        assert bp.ok == ref["ok"]
        assert bp.words == ref["words"] and bp.ali == ref["ali"], (name, i)
        assert bp.tot_bits == ref["tot_bits"], (name, i, bp.tot, ref["tot"])
        assert abs(bp.tot - ref["tot"]) <= 1e-4 * abs(ref["tot"])   # north_star tolerance (met exactly)
        can, cst = _oracle_decode(O, og, cfg, lls[i])
        _compare(bp, dec.frame_stats(i), can, cst, f"{name}/{i}")


def test_no_frames_and_error_paths(oracle_mod):
    from asr_decoder_b200 import _lib
    fst = synth.make_graph(400, 4.0, 20, seed=2)
    g = CudaFst(fst)
    dec = CudaDecoderBatch(g, _cfg(), 2, max_frames=16)
    dec.InitDecoding()
    out = dec.GetBestPath()
    assert [o.ok for o in out] == [False, False] and out[0].status == -7   # nothing decoded yet
    ll = synth.make_loglikes(5, 20, 2.0, seed=1)
    dec.AdvanceDecoding([ll, ll[:0]])                                       # ragged: second stream gets 0 frames
    assert dec.NumFramesDecoded(0) == 5 and dec.NumFramesDecoded(1) == 0
    dec.FinalizeDecoding()
    with pytest.raises(_lib.AsrdError) as e:                                # Advance after Finalize (inl.h:634)
        dec.AdvanceDecoding([ll, ll])
    assert e.value.status == -9
    out = dec.GetBestPath()
    assert out[0].ok and not out[1].ok
    with pytest.raises(_lib.AsrdError):                                     # fewer columns than ilabels
        dec.InitDecoding()
        dec.AdvanceDecoding([ll[:, :10], ll[:, :10]])
    # the same check inside the C ABI (callers that do not go through the Python mirror)
    import ctypes as C
    L = _lib.lib()
    narrow = np.ascontiguousarray(ll[:, :10])
    ptrs = (C.c_void_p * 2)(narrow.ctypes.data, narrow.ctypes.data)
    nfr, strides = (C.c_int32 * 2)(5, 5), (C.c_int32 * 2)(10, 10)
    dec.InitDecoding()
    assert L.asrd_advance_decoding(dec.handles, 2, ptrs, nfr, strides, 10, -1, 0, None) == -1


def test_frames_beyond_max_frames_fail_loudly():
    """More frames than max_frames are never dropped silently: AdvanceDecoding fails with
    ASRD_ERR_FRAMES_OVERFLOW before decoding anything, and the stream stays usable."""
    from asr_decoder_b200 import _lib
    fst = synth.make_graph(400, 4.0, 20, seed=2)
    g = CudaFst(fst)
    dec = CudaDecoderBatch(g, _cfg(), 1, max_frames=16)
    ll = synth.make_loglikes(24, 20, 2.0, seed=1)
    dec.InitDecoding()
    with pytest.raises(_lib.AsrdError) as e:
        dec.AdvanceDecoding([ll])
    assert e.value.status == -6 and dec.NumFramesDecoded(0) == 0
    dec.AdvanceDecoding([ll[:10]])
    with pytest.raises(_lib.AsrdError) as e:           # 10 decoded + 10 more > 16
        dec.AdvanceDecoding([ll[10:20]])
    assert e.value.status == -6 and dec.NumFramesDecoded(0) == 10
    dec.AdvanceDecoding([ll[10:16]])
    dec.FinalizeDecoding()
    assert dec.GetBestPath()[0].ok and dec.NumFramesDecoded(0) == 16


def test_arena_overflow_is_reported_not_silent():
    fst = synth.make_graph(3000, 5.0, 50, seed=4)
    g = CudaFst(fst)
    dec = CudaDecoderBatch(g, _cfg(), 1, max_frames=64, token_capacity=2000)
    ll = synth.make_loglikes(40, 50, 2.0, seed=3)
    out = dec.Decode([ll])
    assert dec.status(0) == -5 and not out[0].ok and out[0].status == -5


@pytest.mark.parametrize("env", [
    {"ASRD_STREAM_KERNEL": "0"},                 # HBM-map kernels only (k_expand + k_post)
    {"ASRD_DEBUG_FLAGS": "8"},                   # on-chip loop, every frame redone through the HBM map
    {"ASRD_DEBUG_FLAGS": str(1500 << 8)},        # on-chip budget of 1500 states: overflow mid-frame
    {"ASRD_STREAM_U": "1"},
])
def test_every_kernel_path_gives_the_same_search(oracle_mod, monkeypatch, env):
    """The on-chip frame loop (k_stream), its mid-frame overflow into the HBM map and the HBM-map
    kernels are three routes through the same search: identical one-best, per-frame cutoffs and
    token counts, all equal to the canonical oracle."""
    O = oracle_mod
    fst = synth.make_graph(60000, 5.0, 500, seed=99)
    lls = [synth.make_loglikes(120, 500, 2.0 + 0.5 * (i % 3), seed=910 + i) for i in range(5)]
    cfg = _cfg()
    g = CudaFst(fst)
    base = CudaDecoderBatch(g, cfg, len(lls), max_frames=128, collect_stats=True)
    out0 = base.Decode(lls)
    st0 = [base.frame_stats(i) for i in range(len(lls))]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    dec = CudaDecoderBatch(g, cfg, len(lls), max_frames=128, collect_stats=True)
    out1 = dec.Decode(lls)
    og = O.OracleGraph(fst)
    for i in range(len(lls)):
        assert dec.status(i) == 0
        st1 = dec.frame_stats(i)
        for f in ("n_in", "n_tokens", "arcs_expanded"):
            assert np.array_equal(st0[i][f], st1[f]), (env, i, f)
        for f in ("cur_cutoff", "abeam", "next_cutoff", "best"):
            assert np.array_equal(st0[i][f].view(np.uint32), st1[f].view(np.uint32)), (env, i, f)
        assert np.array_equal(out0[i].ilabel, out1[i].ilabel) and out0[i].tot_bits == out1[i].tot_bits
        if i < 2:
            ref, rst = _oracle_decode(O, og, cfg, lls[i])
            _compare(out1[i], st1, ref, rst, f"{env} stream {i}")


@pytest.mark.parametrize("case", ["wide-rows", "very-wide-rows", "deep-eps", "tiny-beam"])
def test_on_chip_loop_corner_shapes(oracle_mod, case):
    """Shapes that take the less-travelled branches of the on-chip frame loop: (wide-rows) 6000 pdfs —
    a 24 KB log-likelihood row next to a smaller shared-memory map; (very-wide-rows) 30000 pdfs — the row
    does not fit and is read from global memory; (deep-eps) 60 % of the states have an eps arc to a near state — many closure rounds
    (worklists, three-counter barrier protocol); (tiny-beam) a handful of tokens per frame."""
    O = oracle_mod
    if case == "wide-rows":
        fst, P, T, cfg = synth.make_graph(20000, 5.0, 6000, seed=31), 6000, 60, _cfg()
    elif case == "very-wide-rows":   # 30000 pdfs: the row is read from global memory (k_stream<false>)
        fst, P, T, cfg = synth.make_graph(20000, 5.0, 30000, seed=34), 30000, 40, _cfg()
    elif case == "deep-eps":
        fst, P, T, cfg = synth.make_graph(20000, 4.0, 300, seed=32, p_eps=0.6, eps_span=3), 300, 100, _cfg()
    else:
        fst, P, T, cfg = synth.make_graph(20000, 5.0, 300, seed=33), 300, 150, _cfg(beam=2.0, min_active=0)
    lls = [synth.make_loglikes(T, P, 2.0, seed=500 + i) for i in range(3)]
    g = CudaFst(fst)
    dec = CudaDecoderBatch(g, cfg, len(lls), max_frames=T + 8, collect_stats=True)
    out = dec.Decode(lls)
    og = O.OracleGraph(fst)
    for i, ll in enumerate(lls):
        assert dec.status(i) == 0
        ref, rst = _oracle_decode(O, og, cfg, ll)
        _compare(out[i], dec.frame_stats(i), ref, rst, f"{case} stream {i}")


def test_concurrent_host_threads_on_different_handles(oracle_mod):
    """The reference runs one decoder object per worker thread (v2-asr-service.cc:95-104); the C ABI
    must therefore accept concurrent calls on DIFFERENT handles.  Four host threads decode their own
    batches at the same time on their own CUDA streams (ctypes releases the GIL); every result equals
    the single-threaded one."""
    import threading
    import torch
    fst = synth.make_graph(30000, 5.0, 400, seed=55)
    g = CudaFst(fst)
    cfg = _cfg()
    jobs = [[synth.make_loglikes(80 + 10 * k, 400, 2.0 + 0.25 * (i % 3), seed=700 + 10 * k + i) for i in range(40)]
            for k in range(4)]
    ref = []
    for lls in jobs:
        dec = CudaDecoderBatch(g, cfg, len(lls), max_frames=128)
        ref.append(dec.Decode(lls))
    outs, errs = [None] * len(jobs), []

    def work(k):
        try:
            st = torch.cuda.Stream()
            dec = CudaDecoderBatch(g, cfg, len(jobs[k]), max_frames=128)
            for _ in range(3):
                outs[k] = dec.Decode(jobs[k], stream=st.cuda_stream)
        except Exception as e:  # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(len(jobs))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for k in range(len(jobs)):
        for a, b in zip(ref[k], outs[k]):
            assert a.ok == b.ok and a.tot_bits == b.tot_bits and np.array_equal(a.ilabel, b.ilabel)
