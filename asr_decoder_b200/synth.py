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
