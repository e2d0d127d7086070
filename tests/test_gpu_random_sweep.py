"""Randomised parity sweep: small seeded graphs with dense parallel arcs, eps chains and tight
max-active / min-active settings — the corners where tie-breaking, the trace-back link choice and
the GetCutoff branches matter — decoded by the CUDA path (plain, with arena pruning, and as a CLG
graph) against the canonical oracle: one-best, per-frame statistics and the raw lattice, bit for bit."""
import numpy as np
import pytest

from asr_decoder_b200 import fstio, synth
from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, LatticeFasterDecoderConfig
from test_gpu_lattice import _canon

pytestmark = pytest.mark.gpu


def _check(O, og, dec, lls, cfg, lattice=True):
    for i, ll in enumerate(lls):
        d = O.OracleDecoder(og, O.make_config(cfg.beam, cfg.max_active, cfg.min_active, cfg.lattice_beam), O.MODE_CANONICAL)
        want = d.decode(ll)
        got = dec.GetBestPath(True)[i]
        assert (got.ok, got.words, got.ali) == (want.ok, want.words, want.ali), i
        if want.ok:
            assert got.tot_bits == want.tot_bits, i
        st, ost = dec.frame_stats(i), d.frame_stats()
        assert np.array_equal(st["n_tokens"], ost["n_raw"]), i
        for k in ("cur_cutoff", "next_cutoff", "best"):
            assert st[k].view(np.uint32).tolist() == ost[k].view(np.uint32).tolist(), (i, k)
        if lattice and want.ok:
            lat = dec.GetRawLattice(i)
            otoks, olinks = d.dump_lattice()
            gt, gl = _canon(lat[0], lat[1], True)
            ot, ol = _canon(otoks, olinks, False)
            assert gt == ot and gl == ol, i


@pytest.mark.parametrize("seed", range(16))
def test_random_small_graphs(oracle_mod, seed):
    O = oracle_mod
    rng = np.random.default_rng(1000 + seed)
    n_states = int(rng.integers(30, 400))
    n_pdfs = int(rng.integers(6, 40))
    fst = synth.make_graph(n_states, avg_deg=float(rng.uniform(2.5, 6.0)), n_pdfs=n_pdfs, seed=seed, n_words=50,
                           p_final=float(rng.uniform(0.05, 0.4)), p_eps=float(rng.uniform(0.1, 0.5)),
                           eps_span=int(rng.integers(2, 12)), near_span=int(rng.integers(3, 12)))
    T = int(rng.integers(5, 45))
    lls = [synth.make_loglikes(T, n_pdfs, float(rng.uniform(0.5, 3.0)), seed=77 + seed * 4 + i) for i in range(3)]
    cfg = LatticeFasterDecoderConfig(beam=float(rng.uniform(4.0, 16.0)), max_active=int(rng.integers(5, 300)),
                                     min_active=int(rng.integers(0, 40)), lattice_beam=float(rng.uniform(1.0, 9.0)),
                                     prune_interval=int(rng.integers(3, 12)))
    g = CudaFst(fst)
    og = O.OracleGraph(fst)
    dec = CudaDecoderBatch(g, cfg, len(lls), max_frames=64, collect_stats=True)
    dec.Decode(lls)
    _check(O, og, dec, lls, cfg)
    pr = CudaDecoderBatch(g, cfg, len(lls), max_frames=64, collect_stats=True, prune_tokens=True)
    pr.InitDecoding()
    step = int(rng.integers(1, 9))
    for k in range(0, T, step):
        pr.AdvanceDecoding([ll[k:k + step] for ll in lls])
    pr.FinalizeDecoding()
    _check(O, og, pr, lls, cfg)


@pytest.mark.parametrize("seed", range(8))
def test_random_small_clg_graphs(oracle_mod, tmp_path, seed):
    O = oracle_mod
    rng = np.random.default_rng(2000 + seed)
    n_pdfs = int(rng.integers(8, 40))
    clg, hmms = synth.make_clg(int(rng.integers(20, 200)), n_hmms=int(rng.integers(3, 20)), n_pdfs=n_pdfs,
                               avg_deg=float(rng.uniform(2.5, 5.0)), seed=seed, n_words=40,
                               p_final=float(rng.uniform(0.05, 0.3)), p_eps=float(rng.uniform(0.1, 0.4)))
    gp, hp = str(tmp_path / "c.fst"), str(tmp_path / "h.bin")
    fstio.write_fst(gp, clg)
    fstio.write_hmm_set(hp, hmms)
    T = int(rng.integers(8, 45))
    lls = [synth.make_loglikes(T, n_pdfs, float(rng.uniform(0.5, 2.5)), seed=500 + seed * 4 + i) for i in range(3)]
    cfg = LatticeFasterDecoderConfig(beam=float(rng.uniform(5.0, 15.0)), max_active=int(rng.integers(8, 400)),
                                     min_active=int(rng.integers(0, 30)), lattice_beam=float(rng.uniform(1.0, 8.0)))
    og = O.OracleGraph(None, clg=fstio.materialize_clg(clg, hmms))
    dec = CudaDecoderBatch(CudaFst.ReadClg(gp, hp), cfg, len(lls), max_frames=64, collect_stats=True)
    dec.Decode(lls)
    _check(O, og, dec, lls, cfg)
    if O.have_ref_clg():   # where the compiled reference agrees with itself under two token orders, it agrees with us
        lp = str(tmp_path / "l.bin")
        fstio.write_loglikes(lp, lls)
        kw = dict(beam=cfg.beam, max_active=cfg.max_active, min_active=cfg.min_active, lattice_beam=cfg.lattice_beam)
        r1 = O.run_ref(gp, lp, stats=False, hmm_path=hp, threads=3, hash_ratio=2.0, **kw)[0]
        r2 = O.run_ref(gp, lp, stats=False, hmm_path=hp, threads=3, hash_ratio=1.0, **kw)[0]
        for a, b, x in zip(r1, r2, dec.GetBestPath(True)):
            if (a["words"], a["ali"], a["tot_bits"]) == (b["words"], b["ali"], b["tot_bits"]) and a["ok"]:
                assert x.tot <= a["tot"] * (1 + 1e-6) + 1e-6      # never dearer than a self-stable reference answer
