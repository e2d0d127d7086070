"""Seeded synthetic HCLG graphs and log-likelihood matrices.

The topology follows the survey's generator description: per state, in this order,
an optional final arc to the single super-final state, an optional forward-only
epsilon arc (no epsilon cycles), one emitting self-loop, and ``k`` emitting arcs,
80 % of them to a nearby state.  Epsilon arcs precede emitting arcs in every row
(the layout ``convert_fst`` produces, reference
``src/fst_format_convert_tool/read_fst.c:110-135``).  Parallel arcs ``s -> d`` with
distinct labels do occur and are deliberately kept: they exercise the reference's
trace-back link choice.
"""
from __future__ import annotations

import numpy as np

from .fstio import ARC_DTYPE, Fst


def make_graph(n_states: int, avg_deg: float = 5.0, n_pdfs: int = 200, seed: int = 12345,
               n_words: int = 20000, p_final: float = 0.02, p_eps: float = 0.15,
               eps_span: int = 1000, near_span: int = 64, eps_depth_boost: float = 0.0) -> Fst:
    """Return an ``Fst`` with ``n_states`` real states plus one super-final state."""
    rng = np.random.default_rng(seed)
    S = int(n_states)
    has_final = rng.random(S) < p_final
    has_final[S - 1] = True                     # at least one final, and the tail can end
    eps_dst = np.arange(S, dtype=np.int64) + 1 + rng.integers(0, eps_span + 1, size=S)
    has_eps = (rng.random(S) < p_eps) & (eps_dst < S)
    k = np.rint(avg_deg - 1.17 + rng.uniform(-1.5, 1.5, size=S)).astype(np.int64)
    k = np.maximum(k, 1)
    n_arcs = has_final.astype(np.int64) + has_eps.astype(np.int64) + 1 + k
    off = np.zeros(S + 1, dtype=np.int64)
    np.cumsum(n_arcs, out=off[1:])
    A = int(off[-1])
    arcs = np.zeros(A, dtype=ARC_DTYPE)

    pos = off[:-1].copy()
    # final arcs: 0:0 / U(0,2) -> super-final
    idx = pos[has_final]
    arcs["weight"][idx] = rng.uniform(0.0, 2.0, size=idx.shape[0]).astype(np.float32)
    arcs["nextstate"][idx] = S
    pos += has_final
    # forward epsilon arcs: 0:(word w.p. .5) / U(0,3)
    idx = pos[has_eps]
    ne = idx.shape[0]
    arcs["olabel"][idx] = np.where(rng.random(ne) < 0.5, rng.integers(1, n_words + 1, size=ne), 0)
    arcs["weight"][idx] = rng.uniform(0.0, 3.0, size=ne).astype(np.float32)
    arcs["nextstate"][idx] = eps_dst[has_eps]
    pos += has_eps
    # emitting self-loop: pdf:0 / 0.1+U(0,1)
    arcs["ilabel"][pos] = rng.integers(1, n_pdfs + 1, size=S)
    arcs["weight"][pos] = (0.1 + rng.uniform(0.0, 1.0, size=S)).astype(np.float32)
    arcs["nextstate"][pos] = np.arange(S)
    pos += 1
    # k emitting arcs
    src = np.repeat(np.arange(S, dtype=np.int64), k)
    within = np.arange(src.shape[0], dtype=np.int64) - np.repeat(np.cumsum(k) - k, k)
    idx = np.repeat(pos, k) + within
    m = idx.shape[0]
    near = rng.random(m) < 0.8
    dst = np.where(near, (src + 1 + rng.integers(0, near_span, size=m)) % S, rng.integers(0, S, size=m))
    arcs["ilabel"][idx] = rng.integers(1, n_pdfs + 1, size=m)
    arcs["olabel"][idx] = np.where(rng.random(m) < 0.15, rng.integers(1, n_words + 1, size=m), 0)
    arcs["weight"][idx] = rng.uniform(0.0, 4.0, size=m).astype(np.float32)
    arcs["nextstate"][idx] = dst

    num_arcs = np.concatenate([n_arcs, [0]]).astype(np.uint32)
    niepsilons = np.concatenate([has_final.astype(np.int64) + has_eps.astype(np.int64), [0]]).astype(np.uint32)
    # output-epsilon counts per row (olabel == 0)
    oeps = (arcs["olabel"] == 0).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(oeps)])
    noepsilons = np.concatenate([csum[off[1:]] - csum[off[:-1]], [0]]).astype(np.uint32)
    return Fst(start=0, final_state=S, arcs=arcs, num_arcs=num_arcs,
               niepsilons=niepsilons, noepsilons=noepsilons)


def make_loglikes(n_frames: int, n_pdfs: int, sigma: float = 2.0, seed: int = 777) -> np.ndarray:
    """T x P float32, each row ``log_softmax(N(0, sigma^2))``.
    ``sigma`` selects the regime: 2 = busy (max-active binding), 3 = peaked."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n_frames, n_pdfs)).astype(np.float32) * np.float32(sigma)
    x = x - x.max(axis=1, keepdims=True)
    lse = np.log(np.exp(x.astype(np.float64)).sum(axis=1, keepdims=True))
    return (x.astype(np.float64) - lse).astype(np.float32)


def make_tiny_graph(seed: int = 1) -> Fst:
    """A few dozen states with dense parallel arcs, epsilon chains and ties-free
    weights; used by the pure-Python unit checks."""
    return make_graph(40, avg_deg=4.0, n_pdfs=8, seed=seed, n_words=30, p_final=0.1,
                      p_eps=0.35, eps_span=6, near_span=5)


def make_clg(n_states: int, n_hmms: int = 40, n_pdfs: int = 60, avg_deg: float = 4.0, seed: int = 7,
             n_words: int = 500, p_final: float = 0.05, p_eps: float = 0.2):
    """A seeded CLG graph + HMM set in the shape ``ClgFst`` expects (reference
    ``src/my-decoder/clg-fst.h:9-189``): the CLG graph is a ``make_graph`` graph whose non-eps
    ilabels are HMM ids (1..n_hmms); every HMM is a three-state left-to-right model — state k has a
    self-loop and a forward arc (ilabel = pdf id + 1), the last emitting state ends in an eps arc
    (``_input == 0`` = "hmm end state", clg-fst.h:124-131) — plus an unused final state.
    Returns ``(clg, hmms)``."""
    from .fstio import ARC_DTYPE, Fst
    clg = make_graph(n_states, avg_deg, n_hmms, seed=seed, n_words=n_words, p_final=p_final, p_eps=p_eps,
                     eps_span=max(4, n_states // 10), near_span=max(4, min(64, n_states // 4)))
    rng = np.random.default_rng(seed + 1000)
    hmms = []
    for _ in range(n_hmms):
        rows = []
        for k in range(3):
            row = []
            if k == 2:
                row.append((0, 0, np.float32(rng.uniform(0.0, 0.7)), 3))          # HMM end (eps first)
            row.append((int(rng.integers(1, n_pdfs + 1)), 0, np.float32(rng.uniform(0.1, 1.2)), k))   # self-loop
            if k < 2:
                row.append((int(rng.integers(1, n_pdfs + 1)), 0, np.float32(rng.uniform(0.1, 1.2)), k + 1))
            rows.append(row)
        rows.append([])
        flat = [x for r in rows for x in r]
        arcs = np.zeros(len(flat), dtype=ARC_DTYPE)
        arcs["ilabel"] = [x[0] for x in flat]
        arcs["olabel"] = [x[1] for x in flat]
        arcs["weight"] = [x[2] for x in flat]
        arcs["nextstate"] = [x[3] for x in flat]
        num = np.array([len(r) for r in rows], np.uint32)
        nie = np.array([sum(1 for x in r if x[0] == 0) for r in rows], np.uint32)
        noe = np.array([len(r) for r in rows], np.uint32)
        hmms.append(Fst(0, 3, arcs, num, nie, noe))
    return clg, hmms
