"""Dump cuda vs reference n-best (oracle/_ref/dropin_nbest) for the golden fixtures and a config-3-shaped case."""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asr_decoder_b200 import fstio, synth
BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_nbest")
def run(which, g, l, **cfg):
    cmd = [BIN, f"--graph={g}", f"--loglikes={l}", f"--decoder={which}", "--nbest=10"] + [f"--{k.replace('_','-')}={v}" for k, v in cfg.items()]
    out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    return [json.loads(x) for x in out.splitlines() if x.startswith("{")]
res = {}
for name in ("g1", "g2", "g3"):
    meta = json.load(open(f"{ROOT}/tests/golden/{name}.json"))
    cfg = {k: meta["config"][k] for k in ("beam", "max_active", "min_active", "lattice_beam")}
    g, l = f"{ROOT}/tests/golden/{name}.fst", f"{ROOT}/tests/golden/{name}.llb"
    res[name] = {"cuda": run("cuda", g, l, **cfg), "ref": run("ref", g, l, **cfg)}
tmp = tempfile.mkdtemp()
fst = synth.make_graph(200000, 3.0, 500, seed=777)
lls = [synth.make_loglikes(100, 500, 1.2, seed=50 + i) for i in range(3)]
fstio.write_fst(tmp + "/g.fst", fst); fstio.write_loglikes(tmp + "/l.llb", lls)
res["c3"] = {"cuda": run("cuda", tmp + "/g.fst", tmp + "/l.llb"), "ref": run("ref", tmp + "/g.fst", tmp + "/l.llb")}
for k, v in res.items():
    for side in v.values():
        for u in side:
            u.pop("ali", None)
json.dump(res, open(f"{ROOT}/gpurun_out/nbest_dump.json", "w"))
for k, v in res.items():
    for c, r in zip(v["cuda"], v["ref"]):
        print(k, c["utt"], "raw", c["raw_states"], r["raw_states"], "det", c["det_states"], r["det_states"],
              "nbest cuda", [round(p["tot"], 3) for p in c["nbest"]], "ref", [round(p["tot"], 3) for p in r["nbest"]],
              "same words", [tuple(a["words"]) == tuple(b["words"]) for a, b in zip(c["nbest"], r["nbest"])])
