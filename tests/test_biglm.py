"""biglm (on-the-fly LM-difference composition, BASELINE.json configs[3]).

CPU: the oracle's biglm restatement against the golden vectors of the COMPILED reference
(OnlineLatticeDecoderMempoolBiglm) on the bug-neutral unigram LM pair.  GPU: the CUDA biglm
decoder against the canonical oracle (unigram and bigram LMs) and against the reference golden."""
import json
import os

import numpy as np
import pytest

from asr_decoder_b200 import fstio, lm as LM, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load_b1():
    fst = fstio.read_fst(os.path.join(GOLD, "b1.fst"))
    lls = fstio.read_loglikes(os.path.join(GOLD, "b1.llb"))
    meta = json.load(open(os.path.join(GOLD, "b1.json")))
    lm1 = LM.read_lm(os.path.join(GOLD, "b1.lm1")).Rescale(-1.0)   # kaldi-hclg-my-decoder-biglm.cc:55-60
    lm2 = LM.read_lm(os.path.join(GOLD, "b1.lm2"))
    return fst, lls, meta, lm1, lm2


def test_lm_file_round_trip_and_layout(tmp_path):
    lm = LM.make_lm(30, seed=3, order=2)
    p = str(tmp_path / "x.lm")
    LM.write_lm(p, lm)
    back = LM.read_lm(p)
    assert (back.bos, back.eos, back.unk) == (31, 32, 33)
    assert np.array_equal(back.states, lm.states) and np.array_equal(back.arcs, lm.arcs)
    off = lm.arc_off
    assert np.array_equal(lm.arcs["wordid"][:off[1]], np.arange(off[1]))      # unigram state: direct index
    for s in range(1, len(lm.states)):
        w = lm.arcs["wordid"][off[s]:off[s + 1]]
        assert np.all(np.diff(w) > 0)                                           # sorted for the binary search
    assert (LM.make_lm(30, seed=3, order=1).states["arc_num"][1:] == 0).all()   # unigram-only: empty histories


def test_oracle_biglm_reference_mode_equals_compiled_reference(oracle_mod):
    O = oracle_mod
    fst, lls, meta, lm1, lm2 = _load_b1()
    og, o1, o2 = O.OracleGraph(fst), O.OracleLm(lm1), O.OracleLm(lm2)
    d = O.OracleDecoder(og, O.make_config(**meta["config"]), O.MODE_REFERENCE, o1, o2)
    for ll, ref in zip(lls, meta["reference"]):
        r = d.decode(ll)
        assert r.words == ref["words"] and r.ali == ref["ali"] and r.tot_bits == ref["tot_bits"]
        st = d.frame_stats()
        assert np.array_equal(st["n_raw"], np.array(ref["n_raw"], dtype=np.uint32))
        assert np.array_equal(st["n_within"], np.array(ref["n_within"], dtype=np.uint32))
        assert np.array_equal(st["next_cutoff"].view(np.uint32), np.array(ref["next_cutoff_bits"], dtype=np.uint32))


def test_oracle_biglm_canonical_one_best_equals_reference(oracle_mod):
    O = oracle_mod
    fst, lls, meta, lm1, lm2 = _load_b1()
    og, o1, o2 = O.OracleGraph(fst), O.OracleLm(lm1), O.OracleLm(lm2)
    for i, (ll, ref) in enumerate(zip(lls, meta["reference"])):
        assert meta["self_stable"][i]
        d = O.OracleDecoder(og, O.make_config(**meta["config"]), O.MODE_CANONICAL, o1, o2)
        r = d.decode(ll)
        assert r.words == ref["words"] and r.ali == ref["ali"] and r.tot_bits == ref["tot_bits"]
        # rescoring changes the result: the plain decoder finds another path on this input
        p = O.OracleDecoder(og, O.make_config(**meta["config"]), O.MODE_CANONICAL).decode(ll)
        assert p.tot_bits != r.tot_bits


@pytest.mark.gpu
@pytest.mark.parametrize("order", [1, 2])
def test_cuda_biglm_equals_canonical_oracle(oracle_mod, order):
    from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, CudaLm, LatticeFasterDecoderConfig
    O = oracle_mod
    fst = synth.make_graph(3000, 5.0, 90, seed=41, n_words=60, eps_span=250)
    lm1 = LM.make_lm(60, seed=7, order=order).Rescale(-1.0)
    lm2 = LM.make_lm(60, seed=8, order=order, bigram_density=0.2)
    lls = [synth.make_loglikes(t, 90, s, seed=300 + t) for t, s in ((70, 2.0), (33, 2.5), (1, 2.0), (90, 1.8))]
    cfg = LatticeFasterDecoderConfig(beam=12.0, max_active=2500, min_active=150, lattice_beam=7.0)
    g = CudaFst(fst)
    dec = CudaDecoderBatch(g, cfg, len(lls), max_frames=96, collect_stats=True, old_lm=CudaLm(lm1), new_lm=CudaLm(lm2))
    out = dec.Decode(lls)
    og, o1, o2 = O.OracleGraph(fst), O.OracleLm(lm1), O.OracleLm(lm2)
    for i, ll in enumerate(lls):
        assert dec.status(i) == 0
        d = O.OracleDecoder(og, O.make_config(cfg.beam, cfg.max_active, cfg.min_active, cfg.lattice_beam),
                            O.MODE_CANONICAL, o1, o2)
        ref = d.decode(ll)
        bp = out[i]
        assert bp.ok == ref.ok
        assert bp.words == ref.words and bp.ali == ref.ali, (order, i)
        assert bp.tot_bits == ref.tot_bits, (order, i, bp.tot, ref.tot)
        assert np.array_equal(bp.graph.view(np.uint32), ref.graph.view(np.uint32))
        st, rst = dec.frame_stats(i), d.frame_stats()
        assert np.array_equal(st["n_tokens"], rst["n_raw"]), (order, i)
        assert np.array_equal(st["next_cutoff"].view(np.uint32), rst["next_cutoff"].view(np.uint32))
        assert np.array_equal(st["cur_cutoff"].view(np.uint32), rst["cur_cutoff"].view(np.uint32))
        assert np.array_equal(st["arcs_expanded"].astype(np.int64), rst["arcs_expanded"])


@pytest.mark.gpu
def test_cuda_biglm_equals_compiled_reference_golden():
    from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, CudaLm, LatticeFasterDecoderConfig
    fst, lls, meta, lm1, lm2 = _load_b1()
    cfg = LatticeFasterDecoderConfig(**meta["config"])
    dec = CudaDecoderBatch(CudaFst(fst), cfg, len(lls), max_frames=96, old_lm=CudaLm(lm1), new_lm=CudaLm(lm2))
    # chunked like the streaming service, then once more in one go on the reused decoders
    dec.InitDecoding()
    for f0 in range(0, 80, 30):
        dec.AdvanceDecoding([ll[f0:f0 + 30] for ll in lls])
    dec.FinalizeDecoding()
    for out in (dec.GetBestPath(), dec.Decode(lls)):
        for bp, ref in zip(out, meta["reference"]):
            assert bp.ok and bp.words == ref["words"] and bp.ali == ref["ali"] and bp.tot_bits == ref["tot_bits"]
            assert abs(bp.tot - ref["tot"]) <= 1e-4 * abs(ref["tot"])


@pytest.mark.gpu
@pytest.mark.parametrize("order", [1, 2])
def test_cuda_biglm_raw_lattice_equals_canonical_oracle(oracle_mod, order):
    """GetRawLattice of the biglm decoder (the reference inherits it from the base class,
    online-decoder-base-inl.h:868-975, with the final costs of …-biglm.h:157-215): surviving tokens
    (cost and extra_cost bits) and forward links (labels, graph cost incl. the LM-difference score,
    acoustic cost) equal the canonical oracle's.  Tokens that share an HCLG state are told apart by
    their cost."""
    from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, CudaLm, LatticeFasterDecoderConfig
    O = oracle_mod
    fst = synth.make_graph(3000, 5.0, 90, seed=41, n_words=60, eps_span=250)
    lm1 = LM.make_lm(60, seed=7, order=order).Rescale(-1.0)
    lm2 = LM.make_lm(60, seed=8, order=order, bigram_density=0.2)
    lls = [synth.make_loglikes(t, 90, s, seed=300 + t) for t, s in ((70, 2.0), (33, 2.5))]
    cfg = LatticeFasterDecoderConfig(beam=12.0, max_active=2500, min_active=150, lattice_beam=7.0)
    dec = CudaDecoderBatch(CudaFst(fst), cfg, len(lls), max_frames=96, old_lm=CudaLm(lm1), new_lm=CudaLm(lm2))
    dec.Decode(lls)
    og, o1, o2 = O.OracleGraph(fst), O.OracleLm(lm1), O.OracleLm(lm2)
    bits = lambda x: int(np.float32(x).view(np.uint32))
    for i, ll in enumerate(lls):
        d = O.OracleDecoder(og, O.make_config(cfg.beam, cfg.max_active, cfg.min_active, cfg.lattice_beam),
                            O.MODE_CANONICAL, o1, o2)
        d.decode(ll)
        otoks, olinks = d.dump_lattice()
        toks, links = dec.GetRawLattice(i)
        gt = sorted((int(x["frame"]), int(x["state"]), bits(x["cost"]), bits(x["extra"])) for x in toks)
        ot = sorted((int(x["frame"]), int(x["state"]), bits(x["tot"]), bits(x["extra"])) for x in otoks)
        assert gt == ot, (order, i, len(gt), len(ot))
        gl = sorted((int(toks[x["src"]]["frame"]), int(toks[x["src"]]["state"]), bits(toks[x["src"]]["cost"]),
                     int(toks[x["dst"]]["frame"]), int(toks[x["dst"]]["state"]), bits(toks[x["dst"]]["cost"]),
                     int(x["ilabel"]), int(x["olabel"]), bits(x["graph"]), bits(x["acoustic"])) for x in links)
        ol = sorted((int(x["src_frame"]), int(x["src_state"]), bits(x["src_tot"]), int(x["dst_frame"]),
                     int(x["dst_state"]), bits(x["dst_tot"]), int(x["ilabel"]), int(x["olabel"]), bits(x["graph"]),
                     bits(x["acoustic"])) for x in olinks)
        assert gl == ol, (order, i, len(gl), len(ol))
        assert toks["is_final"].sum() >= 1
