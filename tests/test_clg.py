"""CLG graphs: ClgFst (src/my-decoder/clg-fst.h:9-189) under the reference's
CLG decoder OnlineClgLatticeDecoderMempool (src/my-decoder/online-clg-decoder-mempool-base.h).

CPU: the oracle on the MATERIALISED graph (fstio.materialize_clg: the graph ClgFst expands on the
fly, over the reference's own two-level state ids) in reference-order mode equals the compiled,
unmodified reference CLG decoder bit for bit — one-best, per-frame token counts and cutoffs, with
max-active binding and under three token orders.  That pins the materialisation AND the three
places where the CLG decoder differs from the HCLG one (strict token cutoff, inclusive arc
admission, two-weight pre-pass).  GPU: the CUDA path on the library's own materialisation
(asrd_graph_read_clg) equals the canonical oracle per frame, and the compiled reference's one-best."""
import os

import numpy as np
import pytest

from asr_decoder_b200 import fstio, synth

CASES = [(3, 1.5, 400), (4, 2.0, 300), (5, 1.0, 2000)]   # seed, sigma, max_active
CFG = dict(beam=12.0, min_active=20, lattice_beam=6.0)


def make_case(tmp_path, seed, sigma, n_utts=4, T=70):
    clg, hmms = synth.make_clg(400, n_hmms=25, n_pdfs=50, seed=seed)
    lls = [synth.make_loglikes(T, 50, sigma, seed=10 + i) for i in range(n_utts)]
    gp, hp, lp = (str(tmp_path / x) for x in ("clg.fst", "hmm.bin", "ll.bin"))
    fstio.write_fst(gp, clg)
    fstio.write_hmm_set(hp, hmms)
    fstio.write_loglikes(lp, lls)
    return clg, hmms, lls, gp, hp, lp


def test_hmm_set_round_trip(tmp_path):
    clg, hmms, _, _, hp, _ = make_case(tmp_path, 3, 1.5, n_utts=1)
    back = fstio.read_hmm_set(hp)
    assert len(back) == len(hmms)
    for a, b in zip(hmms, back):
        assert a.arcs.tobytes() == b.arcs.tobytes() and (a.num_arcs == b.num_arcs).all()
    mg = fstio.materialize_clg(clg, hmms)
    assert mg.offset == clg.total_arcs + 1 and mg.fst.eps_first()
    # every emitting CLG arc became the emitting arcs of state 0 of its HMM, two-weight
    n_two = int(mg.from_clg.sum())
    assert n_two == 2 * int((clg.arcs["ilabel"] != 0).sum())        # (state 0 of every synthetic HMM: loop + forward)
    assert np.array_equal((mg.w_hmm + mg.w_clg).astype(np.float32)[mg.from_clg], mg.fst.arcs["weight"][mg.from_clg])


@pytest.mark.parametrize("seed,sigma,max_active", CASES)
def test_oracle_on_the_materialised_graph_equals_the_compiled_clg_reference(oracle_mod, tmp_path, seed, sigma, max_active):
    O = oracle_mod
    if not O.have_ref_clg():
        pytest.skip("oracle/_ref/ref_decode_clg not built on this box")
    clg, hmms, lls, gp, hp, lp = make_case(tmp_path, seed, sigma)
    og = O.OracleGraph(None, clg=fstio.materialize_clg(clg, hmms))
    cfg = dict(CFG, max_active=max_active)
    binding = False
    for hr in (2.0, 1.0, 1.3):
        # (one decoder object per utterance: the reference's HashList keeps its size across utterances)
        res, _ = O.run_ref(gp, lp, stats=True, hmm_path=hp, threads=len(lls), hash_ratio=hr, **cfg)
        for ll, ref in zip(lls, res):
            d = O.OracleDecoder(og, O.make_config(hash_ratio=hr, **cfg), O.MODE_REFERENCE)
            r = d.decode(ll)
            st = d.frame_stats()
            assert (r.ok, r.words, r.ali, r.tot_bits) == (ref["ok"], ref["words"], ref["ali"], ref["tot_bits"])
            assert list(st["n_raw"]) == ref["n_raw"] and list(st["n_in"]) == ref["n_in"]
            for k in ("cur_cutoff", "next_cutoff", "abeam", "best"):
                assert [int(x) for x in st[k].view(np.uint32)] == ref[k + "_bits"], k
            binding |= max(ref["n_raw"]) > max_active
            # the order-independent semantics the CUDA path implements give the same one-best here
            c = O.OracleDecoder(og, O.make_config(**cfg), O.MODE_CANONICAL).decode(ll)
            assert (c.words, c.ali, c.tot_bits) == (ref["words"], ref["ali"], ref["tot_bits"])
    assert binding or max_active >= 2000


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{}, {"ASRD_STREAM_KERNEL": "0"}, {"ASRD_DEBUG_FLAGS": str(300 << 8)}])
@pytest.mark.parametrize("seed,sigma,max_active", CASES)
def test_cuda_clg_equals_canonical_oracle_and_reference(oracle_mod, tmp_path, monkeypatch, seed, sigma, max_active, env):
    """Three routes: the on-chip frame loop (k_stream<.., CLG>), the HBM-map kernels, and the on-chip
    loop with a 300-state budget (frames overflow into the HBM map mid-frame)."""
    from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, LatticeFasterDecoderConfig
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    O = oracle_mod
    clg, hmms, lls, gp, hp, lp = make_case(tmp_path, seed, sigma)
    mg = fstio.materialize_clg(clg, hmms)
    og = O.OracleGraph(None, clg=mg)
    cfg = dict(CFG, max_active=max_active)
    g = CudaFst.ReadClg(gp, hp)
    dec = CudaDecoderBatch(g, LatticeFasterDecoderConfig(**cfg), len(lls), max_frames=80, collect_stats=True)
    out = dec.Decode(lls)
    ref = O.run_ref(gp, lp, stats=False, hmm_path=hp, threads=len(lls), **cfg)[0] if O.have_ref_clg() else None
    for i, ll in enumerate(lls):
        d = O.OracleDecoder(og, O.make_config(**cfg), O.MODE_CANONICAL)
        want = d.decode(ll)
        ost, st = d.frame_stats(), dec.frame_stats(i)
        assert (out[i].ok, out[i].words, out[i].ali, out[i].tot_bits) == (want.ok, want.words, want.ali, want.tot_bits)
        assert np.array_equal(st["n_tokens"], ost["n_raw"])
        for a, b in (("cur_cutoff", "cur_cutoff"), ("next_cutoff", "next_cutoff"), ("best", "best"), ("abeam", "abeam")):
            assert st[a].view(np.uint32).tolist() == ost[b].view(np.uint32).tolist(), a
        assert np.array_equal(st["arcs_expanded"].astype(np.int64), ost["arcs_expanded"].astype(np.int64))
        if ref is not None:
            assert (out[i].words, out[i].ali, out[i].tot_bits) == (ref[i]["words"], ref[i]["ali"], ref[i]["tot_bits"])
        # raw lattice: tokens and links of the canonical oracle
        from test_gpu_lattice import _canon
        toks, links = dec.GetRawLattice(i)
        otoks, olinks = d.dump_lattice()
        gt, gl = _canon(toks, links, True)
        ot, ol = _canon(otoks, olinks, False)
        assert gt == ot and gl == ol          # tokens (cost, extra cost) and links bit-identical
    # ... and with PruneActiveTokens every 20 frames on the device nothing changes
    pr = CudaDecoderBatch(g, LatticeFasterDecoderConfig(prune_interval=20, **cfg), len(lls), max_frames=80, prune_tokens=True)
    pr.InitDecoding()
    for k in range(0, 70, 20):
        pr.AdvanceDecoding([ll[k:k + 20] for ll in lls])
    pr.FinalizeDecoding()
    for a, b in zip(out, pr.GetBestPath(True)):
        assert (a.ok, a.words, a.ali, a.tot_bits) == (b.ok, b.words, b.ali, b.tot_bits)
    for i in range(len(lls)):
        ta, la = dec.GetRawLattice(i)
        tb, lb = pr.GetRawLattice(i)
        assert ta.tobytes() == tb.tobytes() and la.tobytes() == lb.tobytes()
