"""Raw lattice (GetRawLattice + FinalizeDecoding pruning) from the CUDA path against the canonical
oracle: identical surviving tokens (cost and extra_cost bits) and forward links."""
import json
import os

import numpy as np
import pytest

from asr_decoder_b200 import fstio, synth
from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, LatticeFasterDecoderConfig

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle_lattice(O, og, cfg, ll):
    d = O.OracleDecoder(og, O.make_config(cfg.beam, cfg.max_active, cfg.min_active, cfg.lattice_beam),
                        O.MODE_CANONICAL)
    d.decode(ll)
    return d.dump_lattice()


def _canon(toks, links, gpu):
    if gpu:
        t = {(int(x["frame"]), int(x["state"])): (int(x["cost"].view(np.uint32)), int(x["extra"].view(np.uint32)))
             for x in toks}
        l = sorted((int(toks[x["src"]]["frame"]), int(toks[x["src"]]["state"]), int(toks[x["dst"]]["frame"]),
                    int(toks[x["dst"]]["state"]), int(x["ilabel"]), int(x["olabel"]),
                    int(x["graph"].view(np.uint32)), int(x["acoustic"].view(np.uint32))) for x in links)
    else:
        t = {(int(x["frame"]), int(x["state"])): (int(x["tot"].view(np.uint32)), int(x["extra"].view(np.uint32)))
             for x in toks}
        l = sorted((int(x["src_frame"]), int(x["src_state"]), int(x["dst_frame"]), int(x["dst_state"]),
                    int(x["ilabel"]), int(x["olabel"]), int(x["graph"].view(np.uint32)),
                    int(x["acoustic"].view(np.uint32))) for x in links)
    return t, l


@pytest.mark.parametrize("kernel", ["1", "0"])   # k_prune<EMIT> (pull sweep, map in shared memory) / k_lattice (HBM maps)
@pytest.mark.parametrize("name", ["g1", "g2", "g3"])
def test_raw_lattice_equals_canonical_oracle(oracle_mod, monkeypatch, name, kernel):
    monkeypatch.setenv("ASRD_LATTICE_KERNEL", kernel)
    O = oracle_mod
    fst = fstio.read_fst(os.path.join(GOLD, name + ".fst"))
    lls = fstio.read_loglikes(os.path.join(GOLD, name + ".llb"))
    meta = json.load(open(os.path.join(GOLD, name + ".json")))
    cfg = LatticeFasterDecoderConfig(**meta["config"])
    g = CudaFst(fst)
    dec = CudaDecoderBatch(g, cfg, len(lls), max_frames=128)
    dec.Decode(lls)
    og = O.OracleGraph(fst)
    for i, ll in enumerate(lls):
        toks, links = dec.GetRawLattice(i)
        otoks, olinks = _oracle_lattice(O, og, cfg, ll)
        gt, gl = _canon(toks, links, True)
        ot, ol = _canon(otoks, olinks, False)
        assert len(gt) == len(toks)                       # one token per (frame, state)
        assert set(gt) == set(ot), (name, i, len(gt), len(ot))
        assert gt == ot, (name, i)                        # cost and extra_cost bit-identical
        assert gl == ol, (name, i, len(gl), len(ol))
        # the reference's own raw lattice has the same order of magnitude (its token set differs by
        # the order-dependent extras)
        ref = meta["reference"][i]
        assert 0.6 < len(toks) / max(ref["raw_states"], 1) < 1.6
        assert toks["is_final"].sum() >= 1
        assert (toks["frame"][links["dst"]] - toks["frame"][links["src"]] == (links["ilabel"] != 0)).all()


def test_lattice_contains_the_best_path(oracle_mod):
    O = oracle_mod
    fst = synth.make_graph(4000, 5.0, 80, seed=31)
    ll = synth.make_loglikes(90, 80, 2.0, seed=9)
    cfg = LatticeFasterDecoderConfig(beam=12.0, max_active=2500, min_active=100, lattice_beam=6.0)
    dec = CudaDecoderBatch(CudaFst(fst), cfg, 1, max_frames=128)
    bp = dec.Decode([ll])[0]
    toks, links = dec.GetRawLattice(0)
    # cheapest complete path through the lattice == the one-best cost (extra_cost 0 chain)
    zero = toks["extra"] == 0.0
    assert zero[toks["frame"] == 0].any() and zero[(toks["frame"] == ll.shape[0]) & (toks["is_final"] == 1)].any()
    # (bp.tot sums the REPORTED arcs, which with parallel arcs may be dearer than the token's own
    # cost — SURVEY.md Appendix B-4 — so it bounds the final token's cost from above)
    best_final = toks["cost"][(toks["is_final"] == 1)].min()
    assert best_final <= bp.tot + 1e-3
    otoks, olinks = _oracle_lattice(O, O.OracleGraph(fst), cfg, ll)
    assert (len(toks), len(links)) == (len(otoks), len(olinks))


def _same_lattice(a, b):
    if a is None or b is None:
        return a is None and b is None
    return a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()


@pytest.mark.parametrize("route", ["emit", "hbm", "mixed"])
def test_batched_raw_lattice_equals_the_single_calls(monkeypatch, route):
    """asrd_get_raw_lattice_batch: one launch over the streams (one CTA each) returns, stream by
    stream, the bytes of asrd_get_raw_lattice — through the shared-memory pull sweep, through the
    HBM-map sweep, and when only SOME streams have a frame beyond the pull sweep's capacity
    (ASRD_PRUNE_CAP: those alone are redone by k_lattice).  A stream without frames gives None."""
    if route == "hbm":
        monkeypatch.setenv("ASRD_LATTICE_KERNEL", "0")
    if route == "mixed":
        monkeypatch.setenv("ASRD_PRUNE_CAP", "1024")
    fst = synth.make_graph(6000, 4.0, 100, seed=77)
    # flat scores: thousands of tokens per frame (beyond the forced capacity); peaked ones: a few hundred
    lls = [synth.make_loglikes(t, 100, s, seed=400 + t) for t, s in ((60, 1.0), (45, 5.0), (60, 1.2), (30, 5.0), (52, 1.1))]
    cfg = LatticeFasterDecoderConfig(beam=13.0, max_active=3000, min_active=100, lattice_beam=7.0)
    dec = CudaDecoderBatch(CudaFst(fst), cfg, len(lls) + 1, max_frames=64)
    dec.InitDecoding()
    dec.AdvanceDecoding(lls + [np.zeros((0, 100), np.float32)])
    dec.FinalizeDecoding()
    single = [dec.GetRawLattice(i) for i in range(len(lls) + 1)]
    batch = dec.GetRawLatticeBatch()
    assert single[-1] is None and batch[-1] is None
    assert all(x is not None and len(x[0]) > 10 for x in single[:-1])
    if route == "mixed":
        widest = [int(dec.arena_frame_tokens(i).max()) for i in range(len(lls))]
        assert min(widest) < 1024 < max(widest), widest      # both routes inside one call
    for i, (a, b) in enumerate(zip(single, batch)):
        assert _same_lattice(a, b), (route, i)
    assert all(_same_lattice(a, b) for a, b in zip(batch, dec.GetRawLatticeBatch()))   # repeatable
