"""Language-model FSAs for the biglm path (on-the-fly LM-difference composition).

The reference keeps an ARPA LM as an FSA (``src/newlm/arpa2fsa.h:217-480``): per state a sorted
arc array ``{int word; float weight; int to}`` plus ``{float backoff; int backoff_id}``; state 0
is the unigram / start state and is direct-indexed by word id (``arpa2fsa.h:211-214``,
``arpa2fsa.cc:244-262``).  Weights are natural-log probabilities; cost = -weight
(``src/newlm/compose-arpalm.cc:52-70``).  This module builds such FSAs directly, converts ARPA
text files (``convert_arpa``: the library's restatement of ``arpa2fsa``), writes / reads the
reference's binary format
(``ArpaLm::Read`` ``arpa2fsa.h:399-439`` + ``Fsa::Read`` ``arpa2fsa.cc:70-176``), and generates
seeded synthetic unigram / bigram LMs.
"""
from __future__ import annotations

import dataclasses

import numpy as np

LM_STATE_DTYPE = np.dtype([("arc_num", "<i4"), ("backoff_prob", "<f4"), ("backoff_id", "<i4")])
LM_ARC_DTYPE = np.dtype([("wordid", "<i4"), ("weight", "<f4"), ("tostateid", "<i4")])


@dataclasses.dataclass
class LmFsa:
    bos: int
    eos: int
    unk: int
    ngram_counts: list
    states: np.ndarray  # LM_STATE_DTYPE
    arcs: np.ndarray    # LM_ARC_DTYPE, grouped by state

    @property
    def arc_off(self) -> np.ndarray:
        off = np.zeros(len(self.states) + 1, np.int64)
        np.cumsum(self.states["arc_num"], out=off[1:])
        return off

    def Rescale(self, scale: float) -> "LmFsa":
        """``Fsa::Rescale`` (arpa2fsa.cc:264-276): the caller scales the OLD LM by -1
        (``kaldi-hclg-my-decoder-biglm.cc:55-60``)."""
        if scale == 1:
            return self
        st, ar = self.states.copy(), self.arcs.copy()
        ar["weight"] = (ar["weight"] * np.float32(scale)).astype(np.float32)
        st["backoff_prob"] = (st["backoff_prob"] * np.float32(scale)).astype(np.float32)
        return LmFsa(self.bos, self.eos, self.unk, list(self.ngram_counts), st, ar)


def write_lm(path: str, lm: LmFsa) -> None:
    with open(path, "wb") as f:
        f.write(np.array([lm.bos, lm.eos, lm.unk], "<i4").tobytes())
        f.write(np.array([len(lm.ngram_counts)], "<u8").tobytes())      # size_t
        f.write(np.array(lm.ngram_counts, "<i4").tobytes())
        f.write(np.array([len(lm.states)], "<i4").tobytes())
        f.write(np.ascontiguousarray(lm.states).tobytes())
        f.write(np.array([len(lm.arcs)], "<i4").tobytes())
        f.write(np.ascontiguousarray(lm.arcs).tobytes())


def read_lm(path: str) -> LmFsa:
    with open(path, "rb") as f:
        bos, eos, unk = (int(x) for x in np.frombuffer(f.read(12), "<i4"))
        n = int(np.frombuffer(f.read(8), "<u8")[0])
        counts = [int(x) for x in np.frombuffer(f.read(4 * n), "<i4")]
        ns = int(np.frombuffer(f.read(4), "<i4")[0])
        states = np.frombuffer(f.read(12 * ns), LM_STATE_DTYPE).copy()
        na = int(np.frombuffer(f.read(4), "<i4")[0])
        arcs = np.frombuffer(f.read(12 * na), LM_ARC_DTYPE).copy()
    return LmFsa(bos, eos, unk, counts, states, arcs)


def make_lm(n_words: int, seed: int, order: int = 2, bigram_density: float = 0.1) -> LmFsa:
    """Seeded synthetic LM over word ids ``1..n_words`` plus ``<s>``, ``</s>``, ``<unk>`` (ids
    ``n_words+1..n_words+3``).  ``order=1``: unigram-only — every history state has no arcs and
    back-off weight 0, the fixture on which the reference's ``DiffArpaLm::GetArc`` state-argument
    quirk is harmless.  ``order=2``: bigrams with back-off."""
    rng = np.random.default_rng(seed)
    V = n_words + 3
    bos, eos, unk = n_words + 1, n_words + 2, n_words + 3
    uni = rng.normal(0.0, 1.0, V + 1).astype(np.float64)
    uni[0] = -np.inf
    uni[bos] = -np.inf
    logp = (uni - np.log(np.exp(uni[np.isfinite(uni)]).sum())).astype(np.float32)
    states = np.zeros(V + 1, LM_STATE_DTYPE)          # state 0 + one history state per word id
    arcs0 = np.zeros(V + 1, LM_ARC_DTYPE)             # direct-indexed by word id, entry 0 unused
    arcs0["wordid"] = np.arange(V + 1)
    arcs0["weight"] = np.where(np.isfinite(logp), logp, np.float32(-99.0))
    arcs0["weight"][0] = 0
    arcs0["tostateid"] = np.arange(V + 1)             # after word w the history state is w
    arcs0["tostateid"][0] = 0
    states["arc_num"][0] = V + 1
    chunks = [arcs0]
    n_bi = 0
    if order >= 2:
        for w in range(1, V + 1):
            if w == eos:
                continue
            k = rng.binomial(V, bigram_density)
            if k == 0:
                continue
            nxt = np.sort(rng.choice(np.arange(1, V + 1), size=k, replace=False))
            nxt = nxt[nxt != bos]
            if len(nxt) == 0:
                continue
            a = np.zeros(len(nxt), LM_ARC_DTYPE)
            a["wordid"] = nxt
            a["weight"] = (logp[nxt] + rng.uniform(0.2, 1.5, len(nxt))).astype(np.float32).clip(max=-0.01)
            a["tostateid"] = nxt
            chunks.append(a)
            states["arc_num"][w] = len(nxt)
            states["backoff_prob"][w] = np.float32(-rng.uniform(0.05, 0.8))
            n_bi += len(nxt)
    return LmFsa(bos, eos, unk, [V, n_bi] if order >= 2 else [V], states, np.concatenate(chunks))


def convert_arpa(arpa_path: str, wordlist_path: str, out_path: str) -> LmFsa:
    """ARPA text LM -> LM FSA file (``asrd_lm_convert_arpa``: the reference's ``arpa2fsa`` tool,
    ``src/newlm/arpa2fsa.cc:311-739``, byte-identical output); returns the FSA read back."""
    from . import _lib
    _lib.check(_lib.lib().asrd_lm_convert_arpa(arpa_path.encode(), wordlist_path.encode(), out_path.encode()),
               "asrd_lm_convert_arpa")
    return read_lm(out_path)
