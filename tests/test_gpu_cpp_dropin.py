"""The C++ drop-in class (asr_decoder_b200/cpp, `CudaLatticeDecoder : public DecoderItf`) driven
through the reference interface by its test binary, against the compiled reference's golden
vectors."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "asr_decoder_b200", "asrd_cpp_decode")
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("name,extra", [("g1", []), ("g2", ["--chunk=30"]), ("g3", ["--pull"]),
                                        ("g1", ["--pull", "--chunk=7"]),
                                        ("g2", ["--chunk=10", "--prune", "--prune-interval=10"])])
def test_decoder_itf_drop_in_matches_reference_golden(name, extra):
    assert os.path.exists(BIN), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    meta = json.load(open(os.path.join(GOLD, name + ".json")))
    cfg = meta["config"]
    cmd = [BIN, f"--graph={GOLD}/{name}.fst", f"--loglikes={GOLD}/{name}.llb", f"--beam={cfg['beam']}",
           f"--max-active={cfg['max_active']}", f"--min-active={cfg['min_active']}",
           f"--lattice-beam={cfg['lattice_beam']}"] + extra
    out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE).stdout.decode()
    res = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    assert len(res) == len(meta["reference"])
    for r, ref in zip(res, meta["reference"]):
        assert r["ok"] == ref["ok"] and r["frames"] == ref["frames"]
        assert r["words"] == ref["words"] and r["ali"] == ref["ali"]
        assert r["tot_bits"] == ref["tot_bits"] and r["lm_bits"] == ref["lm_bits"]


def test_get_raw_lattice_through_decoder_itf(oracle_mod):
    """DecoderItf::GetRawLattice of the C++ class: topologically sorted Lattice whose state / arc
    counts equal the canonical oracle's surviving tokens / links."""
    import numpy as np
    from asr_decoder_b200 import fstio
    O = oracle_mod
    name = "g3"
    meta = json.load(open(os.path.join(GOLD, name + ".json")))
    cfg = meta["config"]
    cmd = [BIN, f"--graph={GOLD}/{name}.fst", f"--loglikes={GOLD}/{name}.llb", f"--beam={cfg['beam']}",
           f"--max-active={cfg['max_active']}", f"--min-active={cfg['min_active']}",
           f"--lattice-beam={cfg['lattice_beam']}", "--lattice"]
    out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE).stdout.decode()
    res = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    fst = fstio.read_fst(os.path.join(GOLD, name + ".fst"))
    lls = fstio.read_loglikes(os.path.join(GOLD, name + ".llb"))
    og = O.OracleGraph(fst)
    for r, ll in zip(res, lls):
        d = O.OracleDecoder(og, O.make_config(**cfg), O.MODE_CANONICAL)
        d.decode(ll)
        nt, nl = d.counts()
        assert (r["raw_states"], r["raw_arcs"]) == (nt, nl)
        assert r["raw_finals"] >= 1


def test_biglm_through_decoder_itf():
    """The C++ drop-in in its biglm form (constructor (fst, config, oldlm, newlm), LM files in the
    reference's ArpaLm::Read format, old LM rescaled by -1) against the compiled biglm reference."""
    meta = json.load(open(os.path.join(GOLD, "b1.json")))
    cfg = meta["config"]
    cmd = [BIN, f"--graph={GOLD}/b1.fst", f"--loglikes={GOLD}/b1.llb", f"--lm1={GOLD}/b1.lm1", f"--lm2={GOLD}/b1.lm2",
           f"--beam={cfg['beam']}", f"--max-active={cfg['max_active']}", f"--min-active={cfg['min_active']}",
           f"--lattice-beam={cfg['lattice_beam']}", "--chunk=25"]
    out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE).stdout.decode()
    res = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    for r, ref in zip(res, meta["reference"]):
        assert r["ok"] and r["words"] == ref["words"] and r["ali"] == ref["ali"] and r["tot_bits"] == ref["tot_bits"]


def test_biglm_raw_lattice_through_decoder_itf(oracle_mod):
    """DecoderItf::GetRawLattice of the biglm drop-in: Lattice state / arc counts equal the canonical
    biglm oracle's surviving tokens / links on the golden biglm fixture."""
    from asr_decoder_b200 import fstio, lm as LM
    O = oracle_mod
    meta = json.load(open(os.path.join(GOLD, "b1.json")))
    cfg = meta["config"]
    cmd = [BIN, f"--graph={GOLD}/b1.fst", f"--loglikes={GOLD}/b1.llb", f"--lm1={GOLD}/b1.lm1", f"--lm2={GOLD}/b1.lm2",
           f"--beam={cfg['beam']}", f"--max-active={cfg['max_active']}", f"--min-active={cfg['min_active']}",
           f"--lattice-beam={cfg['lattice_beam']}", "--lattice"]
    out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE).stdout.decode()
    res = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    fst = fstio.read_fst(os.path.join(GOLD, "b1.fst"))
    lls = fstio.read_loglikes(os.path.join(GOLD, "b1.llb"))
    lm1 = LM.read_lm(os.path.join(GOLD, "b1.lm1")).Rescale(-1.0)
    lm2 = LM.read_lm(os.path.join(GOLD, "b1.lm2"))
    og, o1, o2 = O.OracleGraph(fst), O.OracleLm(lm1), O.OracleLm(lm2)
    assert len(res) == len(lls)
    for r, ll in zip(res, lls):
        d = O.OracleDecoder(og, O.make_config(**cfg), O.MODE_CANONICAL, o1, o2)
        d.decode(ll)
        nt, nl = d.counts()
        assert (r["raw_states"], r["raw_arcs"]) == (nt, nl)
        assert r["raw_finals"] >= 1


def test_clg_graph_through_decoder_itf(oracle_mod, tmp_path):
    """--graph-type=clg: the C++ class on a CLG graph + HMM set (CudaFst::ReadClg = ClgFst::Init)
    against the compiled reference CLG decoder (ref_decode_clg) — same one-best, in 30-frame chunks."""
    from asr_decoder_b200 import fstio, synth
    O = oracle_mod
    if not O.have_ref_clg():
        pytest.skip("oracle/_ref/ref_decode_clg not built on this box")
    clg, hmms = synth.make_clg(400, n_hmms=25, n_pdfs=50, seed=4)
    lls = [synth.make_loglikes(70, 50, 2.0, seed=10 + i) for i in range(3)]
    gp, hp, lp = (str(tmp_path / x) for x in ("clg.fst", "hmm.bin", "ll.bin"))
    fstio.write_fst(gp, clg)
    fstio.write_hmm_set(hp, hmms)
    fstio.write_loglikes(lp, lls)
    cfg = dict(beam=12.0, max_active=300, min_active=20, lattice_beam=6.0)
    ref = O.run_ref(gp, lp, stats=False, hmm_path=hp, threads=len(lls), **cfg)[0]
    cmd = [BIN, f"--graph={gp}", f"--hmm={hp}", f"--loglikes={lp}", f"--beam={cfg['beam']}",
           f"--max-active={cfg['max_active']}", f"--min-active={cfg['min_active']}",
           f"--lattice-beam={cfg['lattice_beam']}", "--chunk=30"]
    out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE).stdout.decode()
    res = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    assert len(res) == len(ref)
    for r, w in zip(res, ref):
        assert (r["ok"], r["words"], r["ali"], r["tot_bits"]) == (w["ok"], w["words"], w["ali"], w["tot_bits"])
