"""Graph ingestion: OpenFst const files through ConstFst::Read +
Fst(const ConstFst&) — reference src/newfst/const-fst.h:189-221, src/newfst/optimize-fst.h:82-134.
CPU: the host-side conversion equals the reference's own (oracle/_ref/const2flat = the unmodified
reference reader + converter, dumped through the Fst accessors) byte for byte.  GPU: the library's
reader (asrd_graph_read_const) yields the same search as the flat file."""
import os
import subprocess

import numpy as np
import pytest

from asr_decoder_b200 import fstio, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "oracle", "_ref", "const2flat")


@pytest.mark.parametrize("seed,p_final,p_eps", [(7, 0.02, 0.15), (8, 0.5, 0.0), (9, 0.0, 0.4)])
def test_const_round_trip_and_reference_conversion(tmp_path, seed, p_final, p_eps):
    fst = synth.make_graph(3000, 5.0, 60, seed=seed, p_final=p_final, p_eps=p_eps)
    cpath, mine, ref = str(tmp_path / "g.const"), str(tmp_path / "mine.fst"), str(tmp_path / "ref.fst")
    fstio.write_const_fst(cpath, fst)
    got = fstio.read_const_fst(cpath)
    assert got.start == fst.start and got.final_state == fst.final_state
    for f in ("arcs", "num_arcs", "niepsilons", "noepsilons"):
        assert np.array_equal(getattr(got, f), getattr(fst, f)), f
    assert got.eps_first()           # the final arc leads its row: what the eps-first kernels rely on
    if not os.path.exists(TOOL):
        pytest.skip("oracle/_ref/const2flat not built on this box")
    fstio.write_fst(mine, got)
    subprocess.check_call([TOOL, cpath, ref], stderr=subprocess.DEVNULL)
    assert open(mine, "rb").read() == open(ref, "rb").read()


def test_const_reader_rejects_other_files(tmp_path):
    p = str(tmp_path / "bad")
    open(p, "wb").write(b"\0" * 64)
    with pytest.raises(IOError):
        fstio.read_const_fst(p)
    fst = synth.make_graph(50, 4.0, 8, seed=1)
    fstio.write_fst(p, fst)          # a flat newfst file is not a const FST
    with pytest.raises(IOError):
        fstio.read_const_fst(p)


@pytest.mark.gpu
def test_library_reads_const_files(tmp_path):
    from asr_decoder_b200 import _lib
    from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, LatticeFasterDecoderConfig
    fst = synth.make_graph(20000, 5.0, 200, seed=3)
    cpath = str(tmp_path / "g.const")
    fstio.write_const_fst(cpath, fst)
    cfg = LatticeFasterDecoderConfig(beam=13.0, max_active=3000, min_active=200, lattice_beam=8.0)
    lls = [synth.make_loglikes(80, 200, 2.0, seed=60 + i) for i in range(3)]
    a = CudaDecoderBatch(CudaFst(fst), cfg, 3, max_frames=96).Decode(lls)
    b = CudaDecoderBatch(CudaFst.ReadConstFst(cpath), cfg, 3, max_frames=96).Decode(lls)
    for x, y in zip(a, b):
        assert x.ok and y.ok and x.words == y.words and x.ali == y.ali and x.tot_bits == y.tot_bits
    bad = str(tmp_path / "bad")
    fstio.write_fst(bad, fst)
    with pytest.raises(_lib.AsrdError) as e:
        CudaFst.ReadConstFst(bad)
    assert e.value.status == -10
