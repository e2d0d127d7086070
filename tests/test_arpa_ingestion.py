"""LM ingestion: ARPA text -> LM FSA file.  The library's converter
(asrd_lm_convert_arpa, asr_decoder_b200/csrc/asrd_arpa.cc — host code, no device) against the
UNMODIFIED reference converter (src/newlm/arpa2fsa.cc driven like arpa2fsa-bin.cc, built into
oracle/_ref/arpa2fsa): byte-identical output files on seeded trigram LMs with the things real ARPA
files have — missing back-off columns, comment and blank lines, -99 for <s>, an out-of-vocabulary
line, trigrams whose suffix bigram does not exist (the back-off search has to shorten twice)."""
import os
import subprocess

import numpy as np
import pytest

from asr_decoder_b200 import lm as LM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "oracle", "_ref", "arpa2fsa")


def write_case(tmp_path, seed, n_words=40, order=3):
    rng = np.random.default_rng(seed)
    words = ["w%d" % i for i in range(n_words)]
    vocab = ["<eps>"] + words + ["<s>", "</s>", "<unk>"]
    ids = {w: i for i, w in enumerate(vocab)}
    wpath = str(tmp_path / "words.txt")
    with open(wpath, "w") as f:
        for w in vocab:
            f.write("%s %d\n" % (w, ids[w]))
    lines1 = []
    for w in words + ["</s>"]:
        lines1.append("%.6f\t%s\t%.6f" % (-rng.uniform(0.5, 4.0), w, -rng.uniform(0.0, 1.0)))
    lines1.insert(3, "-99\t<s>\t%.6f" % -rng.uniform(0.1, 0.9))
    lines1.insert(7, "-2.5\t<unk>\t-0.3")            # dropped by the reference: its id is the unk id
    lines1.insert(9, "-1.5\tnot_in_the_list\t-0.2")  # dropped: unknown word
    if seed % 2:
        lines1[5] = lines1[5].rsplit("\t", 1)[0]     # a unigram without a back-off column
    hist1 = [w for w in ["<s>"] + words if rng.random() < 0.5]
    bigrams = {}
    lines2 = []
    for h in hist1:                                  # grouped by history, as SRILM writes them
        nxt = sorted(set(rng.choice(words + ["</s>"], size=rng.integers(1, 6)).tolist()))
        rng.shuffle(nxt)                             # (inside a group the order is free)
        for w in nxt:
            bigrams[(h, w)] = True
            bo = "" if (w == "</s>" or rng.random() < 0.3 or order == 2) else "\t%.6f" % -rng.uniform(0.0, 0.7)
            lines2.append("%.6f\t%s %s%s" % (-rng.uniform(0.2, 3.0), h, w, bo))
    lines3 = []
    if order >= 3:
        for (h, w) in bigrams:                       # histories in the order their states were made
            if w == "</s>" or rng.random() < 0.5:
                continue
            for c in sorted(set(rng.choice(words + ["</s>"], size=rng.integers(1, 4)).tolist())):
                lines3.append("%.6f\t%s %s %s" % (-rng.uniform(0.2, 3.0), h, w, c))
    apath = str(tmp_path / "lm.arpa")
    with open(apath, "w") as f:
        f.write("# a comment line\n\n\\data\\\n")
        f.write("ngram 1=%d\nngram 2=%d\n" % (len(lines1), len(lines2)))
        if order >= 3:
            f.write("ngram 3=%d\n" % len(lines3))
        f.write("\n\\1-grams:\n" + "\n".join(lines1) + "\n\n\\2-grams:\n" + "\n".join(lines2) + "\n")
        if order >= 3:
            f.write("\n\\3-grams:\n" + "\n".join(lines3) + "\n")
        f.write("\n\\end\\\n")
    return apath, wpath, ids


@pytest.mark.parametrize("seed,order", [(1, 3), (2, 3), (3, 2), (4, 3)])
def test_conversion_is_byte_identical_to_the_reference_tool(tmp_path, seed, order):
    apath, wpath, ids = write_case(tmp_path, seed, order=order)
    mine, ref = str(tmp_path / "mine.fsa"), str(tmp_path / "ref.fsa")
    fsa = LM.convert_arpa(apath, wpath, mine)
    # what the file says, independent of the reference: header, direct-indexed unigram state,
    # sorted arcs, back-off pointers to shorter histories
    assert (fsa.bos, fsa.eos, fsa.unk) == (ids["<s>"], ids["</s>"], ids["<unk>"])
    assert len(fsa.ngram_counts) == order
    off = fsa.arc_off
    a0 = fsa.arcs[off[0]:off[1]]
    assert (a0["wordid"] == np.arange(len(a0))).all() and (a0["tostateid"] == np.arange(len(a0)) + 1).all()
    for s in range(len(fsa.states)):
        w = fsa.arcs["wordid"][off[s]:off[s + 1]]
        assert (np.diff(w) > 0).all()
    assert (fsa.states["backoff_id"] < np.arange(len(fsa.states)).clip(min=1)).all()
    assert fsa.arcs["weight"][off[0] + ids["<s>"]] == np.float32(np.float64(np.float32(-99.0)) * np.log(10.0))
    if not os.path.exists(TOOL):
        pytest.skip("oracle/_ref/arpa2fsa not built on this box")
    subprocess.run([TOOL, apath, wpath, ref], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert open(mine, "rb").read() == open(ref, "rb").read()


def test_histories_out_of_order_are_reported(tmp_path):
    """The reference asserts that a state's arcs stay contiguous (n-grams grouped by history,
    arpa2fsa.h:158-171) and aborts; the library returns an error."""
    from asr_decoder_b200 import _lib
    wpath = str(tmp_path / "w.txt")
    open(wpath, "w").write("<eps> 0\na 1\nb 2\n<s> 3\n</s> 4\n<unk> 5\n")
    apath = str(tmp_path / "bad.arpa")
    open(apath, "w").write("\\data\\\nngram 1=4\nngram 2=3\n\n\\1-grams:\n-1 a -0.1\n-1 b -0.1\n-99 <s> -0.1\n-1 </s>\n\n"
                           "\\2-grams:\n-0.5 a b\n-0.5 b a\n-0.5 a a\n\n\\end\\\n")
    with pytest.raises(_lib.AsrdError):
        LM.convert_arpa(apath, wpath, str(tmp_path / "o.fsa"))
