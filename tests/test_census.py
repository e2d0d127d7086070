"""Parity census (tests/golden/census.json, made by tests/golden/make_census.py): the
order-independent semantics the CUDA library implements against the COMPILED REFERENCE under
seven token visiting orders, on the headline workloads.  CPU part: the committed file is
self-consistent and says what DESIGN.md says; where oracle/_ref exists a sample is re-derived live.
The GPU part (test_gpu_census.py) checks the CUDA output against the canonical digests."""
import hashlib
import itertools
import json
import os

import numpy as np
import pytest

from asr_decoder_b200 import fstio, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CLASSES = ("identical", "one_order", "lower_cost", "higher_cost")

# what DESIGN.md section 5.1 states (counts per class of the committed census)
PINNED = {
    "c2_s2.0": dict(identical=28, one_order=31, lower_cost=1, higher_cost=4),
    "c2_s3.0": dict(identical=46, one_order=14, lower_cost=3, higher_cost=1),
    "c1_s1.5": dict(identical=8, one_order=0, lower_cost=0, higher_cost=0),
    "c1_s2.0": dict(identical=8, one_order=0, lower_cost=0, higher_cost=0),
    "c1_s3.0": dict(identical=7, one_order=1, lower_cost=0, higher_cost=0),
}


def census():
    return json.load(open(os.path.join(GOLD, "census.json")))


def digest(xs) -> str:
    return hashlib.sha1(np.asarray(xs, dtype="<i4").tobytes()).hexdigest()[:16]


def f32(bits):
    return float(np.array([bits], dtype=np.uint32).view(np.float32)[0])


def key(x):
    return (x["words_sha"], x["ali_sha"], x["tot_bits"])


def classify(canon, refs):
    same = [key(r) == key(canon) for r in refs]
    if all(same):
        return "identical"
    if any(same):
        return "one_order"
    return "lower_cost" if f32(canon["tot_bits"]) <= min(f32(r["tot_bits"]) for r in refs) else "higher_cost"


def test_census_classes_are_what_the_docs_state():
    c = census()
    assert len(c["hash_ratios"]) >= 6
    for name, s in c["sets"].items():
        got = {k: 0 for k in CLASSES}
        for u in s["utts"]:
            assert classify(u["canonical"], u["reference"]) == u["class"], (name, u["seed"])
            assert len(u["reference"]) == len(c["hash_ratios"])
            got[u["class"]] += 1
        assert got == PINNED[name], (name, got)
        assert got == s["summary"]["classes"]
    assert sum(len(c["sets"][k]["utts"]) for k in ("c2_s2.0", "c2_s3.0")) >= 128


def test_canonical_is_as_close_to_a_reference_order_as_the_reference_is_to_itself():
    """The honest parity statement on config 2: the reference's one-best
    depends on its token visiting order, and the canonical answer agrees with any one reference
    order about as often as two reference orders agree with each other.  Config 1 (the reference's
    own CPU-sized case) is bit-identical throughout except one sigma=3 utterance that the
    reference itself answers in two ways."""
    c = census()
    for name, s in c["sets"].items():
        U = s["utts"]
        n_orders = len(c["hash_ratios"])
        pair = [np.mean([key(u["reference"][i]) == key(u["reference"][j]) for u in U])
                for i, j in itertools.combinations(range(n_orders), 2)]
        can = [np.mean([key(u["reference"][i]) == key(u["canonical"]) for u in U]) for i in range(n_orders)]
        assert np.mean(can) >= np.mean(pair) - 0.10, (name, np.mean(can), np.mean(pair))
        assert max(can) >= min(pair), name
        # a dearer path than the reference's cheapest answer is the exception, never the rule
        dearer = sum(u["class"] == "higher_cost" for u in U) / len(U)
        assert dearer <= 0.07, (name, dearer)
        for u in U:
            if u["class"] == "higher_cost":   # and never by more than a beam-delta-sized margin per 100 frames
                assert u["tot"] - min(u["reference_tots"]) < 0.01 * u["tot"], (name, u["seed"])
    for name in ("c1_s1.5", "c1_s2.0"):
        assert all(u["class"] == "identical" for u in c["sets"][name]["utts"])


def test_token_count_deviation_is_reported_honestly():
    """Per-frame within-cutoff token counts against the reference (hash ratio 2.0): within 1 % where
    max-active does not bind (config 1 sigma 1.5 / 2), 10-15 % mean per-frame deviation on config 2
    where the reference's order-dependent extras leak into GetCutoff (inl.h:169-203,330-333) —
    while the totals over an utterance agree to 1 %."""
    c = census()
    s = c["sets"]
    assert s["c1_s2.0"]["summary"]["n_within_mean_rel_diff"] < 0.01
    assert s["c1_s1.5"]["summary"]["n_within_mean_rel_diff"] < 0.01
    assert 0.05 < s["c2_s2.0"]["summary"]["n_within_mean_rel_diff"] < 0.20
    assert s["c2_s3.0"]["summary"]["n_within_mean_rel_diff"] < 0.15
    for name in s:
        assert abs(s[name]["summary"]["n_within_total_ratio"] - 1.0) < 0.02, name


@pytest.mark.parametrize("name,n_utts,orders", [("c1_s2.0", 2, (2.0, 1.0, 1.3)), ("c1_s3.0", 8, (2.0, 1.7)),
                                                ("c2_s2.0", 6, (2.0, 1.0))])
def test_live_sample_against_the_compiled_reference(oracle_mod, tmp_path, name, n_utts, orders):
    """Re-derives part of the census where oracle/_ref exists: the committed reference digests are
    what the compiled reference prints today, and the canonical digests are what the oracle's
    canonical mode computes."""
    O = oracle_mod
    if not O.have_ref():
        pytest.skip("oracle/_ref/ref_decode not present on this box")
    c = census()
    s = c["sets"][name]
    g = s["graph"]
    fst = synth.make_graph(g["states"], g["avg_deg"], g["pdfs"], seed=g["seed"])
    utts = s["utts"][:n_utts]
    lls = [synth.make_loglikes(s["frames"], g["pdfs"], s["sigma"], seed=u["seed"]) for u in utts]
    gp, lp = str(tmp_path / "g.fst"), str(tmp_path / "l.llb")
    fstio.write_fst(gp, fst)
    fstio.write_loglikes(lp, lls)
    for hr in orders:
        # the census ran one decoder object per thread over 64 utterances dealt round-robin; the
        # reference's HashList keeps its grown size across utterances, so only utterances decoded
        # FIRST by their thread are reproducible from a shorter list: use as many threads as
        # utterances
        res, _ = O.run_ref(gp, lp, stats=False, threads=n_utts, hash_ratio=hr, **c["config"])
        k = c["hash_ratios"].index(hr)
        for u, r in zip(utts, res):
            if u["seed"] - s["utts"][0]["seed"] >= c["threads"]:
                continue
            want = u["reference"][k]
            assert (digest(r["words"]), digest(r["ali"]), r["tot_bits"]) == key(want), (name, u["seed"], hr)
    og = O.OracleGraph(fst)
    cfg = O.make_config(**c["config"])
    for u, ll in zip(utts, lls):
        r = O.OracleDecoder(og, cfg, O.MODE_CANONICAL).decode(ll)
        assert (digest(r.words), digest(r.ali), r.tot_bits) == key(u["canonical"]), (name, u["seed"])
