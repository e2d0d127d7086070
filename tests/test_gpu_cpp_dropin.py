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
                                        ("g1", ["--pull", "--chunk=7"])])
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
