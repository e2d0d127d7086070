#!/usr/bin/env python
"""Regenerates the golden fixtures from the COMPILED REFERENCE (oracle/_ref/ref_decode, built from
/root/reference by oracle/Makefile).  Run in the build container only (the reference sources do
not exist on the GPU box):

    python tests/golden/make_golden.py

For every case it writes <name>.fst / <name>.llb (the exact inputs) and <name>.json: the
reference's own one-best (words, alignment, cost bits) and per-frame statistics under three token
orders (--hash-ratio 2.0, 1.0, 1.1).  `self_stable` is true when all three orders agree on the
one-best: only then is bit-identity with an order-independent decoder a
meaningful assertion.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from asr_decoder_b200 import fstio, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {
    # name: (graph kwargs, [(frames, sigma, seed)], decoder config)
    "g1": (dict(n_states=300, avg_deg=4.0, n_pdfs=20, seed=11, n_words=50, eps_span=40, near_span=16),
           [(40, 2.0, 1), (25, 3.0, 2), (1, 2.0, 3)],
           dict(beam=13.0, max_active=7000, min_active=200, lattice_beam=8.0)),
    "g2": (dict(n_states=2000, avg_deg=5.0, n_pdfs=100, seed=12, n_words=500, eps_span=200),
           [(80, 2.0, 4), (60, 1.5, 5)],
           dict(beam=13.0, max_active=300, min_active=20, lattice_beam=8.0)),
    "g3": (dict(n_states=2000, avg_deg=5.0, n_pdfs=100, seed=13, n_words=500, eps_span=200, p_eps=0.3),
           [(50, 2.5, 6), (70, 2.0, 7)],
           dict(beam=9.0, max_active=1500, min_active=400, lattice_beam=6.0)),
}


BIGLM_CASES = {
    # unigram-only LM pair: the fixture on which the reference's DiffArpaLm::GetArc state-argument
    # quirk is harmless, so bit-identity with the reference is meaningful
    "b1": (dict(n_states=2000, avg_deg=5.0, n_pdfs=100, seed=12, n_words=50, eps_span=200),
           [(80, 2.0, 4), (40, 2.5, 5)], dict(beam=13.0, max_active=2000, min_active=200, lattice_beam=8.0),
           (50, 1, 2, 1)),   # (n_words, lm1 seed, lm2 seed, order)
}


def make_biglm():
    from asr_decoder_b200 import lm as LM
    for name, (gkw, utts, cfg, (nw, s1, s2, order)) in BIGLM_CASES.items():
        fst = synth.make_graph(**gkw)
        lls = [synth.make_loglikes(t, gkw["n_pdfs"], sig, seed=sd) for (t, sig, sd) in utts]
        gp, lp = os.path.join(HERE, name + ".fst"), os.path.join(HERE, name + ".llb")
        l1p, l2p = os.path.join(HERE, name + ".lm1"), os.path.join(HERE, name + ".lm2")
        fstio.write_fst(gp, fst)
        fstio.write_loglikes(lp, lls)
        LM.write_lm(l1p, LM.make_lm(nw, seed=s1, order=order))
        LM.write_lm(l2p, LM.make_lm(nw, seed=s2, order=order))
        runs = {str(hr): O.run_ref_biglm(gp, lp, l1p, l2p, hash_ratio=hr, **cfg) for hr in (2.0, 1.0, 1.1)}
        base = runs["2.0"]
        stable = [all(runs[k][i]["words"] == base[i]["words"] and runs[k][i]["ali"] == base[i]["ali"] and
                      runs[k][i]["tot_bits"] == base[i]["tot_bits"] for k in runs) for i in range(len(base))]
        out = {"generator": "oracle/_ref/ref_decode_biglm (compiled reference OnlineLatticeDecoderMempoolBiglm)",
               "config": cfg, "graph": gkw, "utts": utts, "lm": {"n_words": nw, "seeds": [s1, s2], "order": order},
               "self_stable": stable, "reference": base,
               "other_orders": {k: v for k, v in runs.items() if k != "2.0"}}
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(out, f)
        print(name, "self_stable", stable, "tot", [r["tot"] for r in base])


def main():
    if O.have_ref_biglm():
        make_biglm()
    if not O.have_ref():
        raise SystemExit("oracle/_ref/ref_decode missing: run `make -C oracle ref` where /root/reference exists")
    for name, (gkw, utts, cfg) in CASES.items():
        fst = synth.make_graph(**gkw)
        lls = [synth.make_loglikes(t, gkw["n_pdfs"], sig, seed=sd) for (t, sig, sd) in utts]
        gp, lp = os.path.join(HERE, name + ".fst"), os.path.join(HERE, name + ".llb")
        fstio.write_fst(gp, fst)
        fstio.write_loglikes(lp, lls)
        runs = {}
        for hr in (2.0, 1.0, 1.1):
            res, _ = O.run_ref(gp, lp, stats=True, lattice=(hr == 2.0), hash_ratio=hr, **cfg)
            for r in res:
                r.pop("seconds", None)
            runs[str(hr)] = res
        base = runs["2.0"]
        stable = [all(runs[k][i]["words"] == base[i]["words"] and runs[k][i]["ali"] == base[i]["ali"] and
                      runs[k][i]["tot_bits"] == base[i]["tot_bits"] for k in runs) for i in range(len(base))]
        out = {"generator": "oracle/_ref/ref_decode (compiled reference OnlineLatticeDecoderMempool)",
               "config": cfg, "graph": gkw, "utts": utts, "self_stable": stable,
               "reference": base,
               "other_orders": {k: [{"words": r["words"], "ali": r["ali"], "tot_bits": r["tot_bits"],
                                     "n_raw": r["n_raw"], "n_within": r["n_within"]} for r in v]
                                for k, v in runs.items() if k != "2.0"}}
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(out, f)
        print(name, "states", fst.total_states, "arcs", fst.total_arcs, "self_stable", stable,
              "tot", [r["tot"] for r in base])


if __name__ == "__main__":
    main()
