"""Host-side logic: file formats, generators, config mirror, sharding.  CPU only."""
import numpy as np
import pytest

from asr_decoder_b200 import fstio, sharding, synth
from asr_decoder_b200.decoder import LatticeFasterDecoderConfig


def test_fst_file_round_trip(tmp_path):
    fst = synth.make_graph(500, 5.0, 30, seed=3)
    p = str(tmp_path / "g.fst")
    fstio.write_fst(p, fst)
    g2 = fstio.read_fst(p)
    assert g2.start == fst.start and g2.final_state == fst.final_state
    assert np.array_equal(g2.arcs, fst.arcs) and np.array_equal(g2.num_arcs, fst.num_arcs)
    assert np.array_equal(g2.niepsilons, fst.niepsilons)
    # header layout of Fst::ReadFst (optimize-fst.h:226-239)
    hdr = np.fromfile(p, dtype="<i4", count=6)
    assert list(hdr[:4]) == [fst.start, fst.final_state, fst.total_states, fst.total_arcs]
    with open(p, "r+b") as f:
        f.truncate(100)
    with pytest.raises(IOError):
        fstio.read_fst(p)


def test_loglikes_round_trip(tmp_path):
    lls = [synth.make_loglikes(t, 17, 2.0, seed=t) for t in (5, 1, 9)]
    p = str(tmp_path / "l.llb")
    fstio.write_loglikes(p, lls)
    back = fstio.read_loglikes(p)
    assert all(np.array_equal(a, b) for a, b in zip(lls, back))
    assert np.allclose(np.exp(lls[0].astype(np.float64)).sum(axis=1), 1.0, atol=1e-5)


def test_generator_invariants():
    fst = synth.make_graph(3000, 5.0, 50, seed=9)
    assert fst.eps_first()                                   # eps arcs first in every row
    assert fst.final_state == fst.total_states - 1           # single super-final state, no arcs
    assert fst.num_arcs[fst.final_state] == 0
    il, ns = fst.arcs["ilabel"], fst.arcs["nextstate"]
    assert il.min() >= 0 and il.max() <= 50
    assert ns.min() >= 0 and ns.max() <= fst.final_state
    off = fst.row_off
    src = np.repeat(np.arange(fst.total_states), fst.num_arcs.astype(np.int64))
    eps = il == 0
    assert np.all(ns[eps] > src[eps])                        # forward-only eps arcs: no eps cycles
    assert (ns == fst.final_state).sum() >= 1                # a reachable final
    again = synth.make_graph(3000, 5.0, 50, seed=9)
    assert np.array_equal(again.arcs, fst.arcs)              # seeded
    assert off[-1] == fst.total_arcs


def test_config_mirror_defaults_and_check():
    c = LatticeFasterDecoderConfig()
    # lattice-faster-decoder-conf.h:35-44
    assert (c.beam, c.min_active, c.lattice_beam, c.prune_interval, c.beam_delta, c.hash_ratio, c.prune_scale) == \
           (16.0, 200, 10.0, 25, 0.5, 2.0, 0.1)
    assert c.max_active == 2 ** 31 - 1
    c.Check()
    with pytest.raises(AssertionError):
        LatticeFasterDecoderConfig(beam=0.0).Check()
    with pytest.raises(AssertionError):
        LatticeFasterDecoderConfig(max_active=1).Check()


def test_stream_sharding_partition():
    for n, w in ((10, 3), (256, 8), (5, 8), (0, 2)):
        parts = [sharding.shard_indices(n, r, w) for r in range(w)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(n))                          # every stream exactly once
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        assert all(p == [i for i in range(n) if i % w == r] for r, p in enumerate(parts))
    merged = sharding.merge_results(7, 2, [["a0", "a2", "a4", "a6"], ["a1", "a3", "a5"]])
    assert merged == [f"a{i}" for i in range(7)]
