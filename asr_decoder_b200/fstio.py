"""On-disk formats shared by the CUDA decoder, the oracle and the reference harness.

* newfst graph file — the reference's flat HCLG format, read by ``Fst::ReadFst``
  (reference ``src/newfst/optimize-fst.h:226-280``):
  six int32 ``start, final_state, total_states, total_arcs, total_niepsilons,
  total_noepsilons``; ``total_states x {u32 num_arcs, niepsilons, noepsilons}``;
  ``total_arcs x {i32 ilabel, i32 olabel, f32 weight, i32 nextstate}``.
* loglikes file — ours: int32 magic ``0x4c4c5341``, int32 n_utt, then per
  utterance int32 T, int32 P, float32[T*P]; column = ilabel - 1, the convention of
  Kaldi's ``DecodableMatrixScaled`` that the reference's bins use
  (``src/kaldi-nnet3bin/kaldi-hclg-my-decoder.cc:107``).
"""
from __future__ import annotations

import dataclasses
import numpy as np

ARC_DTYPE = np.dtype([("ilabel", "<i4"), ("olabel", "<i4"), ("weight", "<f4"), ("nextstate", "<i4")])
STATEINFO_DTYPE = np.dtype([("num_arcs", "<u4"), ("niepsilons", "<u4"), ("noepsilons", "<u4")])
LL_MAGIC = 0x4C4C5341


@dataclasses.dataclass
class Fst:
    """Host mirror of the reference ``Fst`` (``src/newfst/optimize-fst.h:53-307``):
    two flat arrays plus start / super-final ids.  ``arcs`` is the 16-byte record
    array verbatim; ``row_off`` is the exclusive prefix sum of ``num_arcs``."""

    start: int
    final_state: int
    arcs: np.ndarray        # ARC_DTYPE [A]
    num_arcs: np.ndarray    # u32 [S]
    niepsilons: np.ndarray  # u32 [S]
    noepsilons: np.ndarray  # u32 [S]

    @property
    def total_states(self) -> int:
        return int(self.num_arcs.shape[0])

    @property
    def total_arcs(self) -> int:
        return int(self.arcs.shape[0])

    @property
    def row_off(self) -> np.ndarray:
        off = np.zeros(self.total_states + 1, dtype=np.int64)
        np.cumsum(self.num_arcs, out=off[1:])
        return off

    # Fst::Start / IsFinal / NumInputEpsilons (optimize-fst.h:169-192)
    def Start(self) -> int:
        return self.start

    def IsFinal(self, s: int) -> bool:
        return s == self.final_state

    def NumInputEpsilons(self, s: int) -> int:
        return int(self.niepsilons[s])

    def eps_first(self) -> bool:
        """True when every state's input-epsilon arcs form a prefix of its row — the
        layout ``convert_fst`` / ``Fst(const ConstFst&)`` produce and the layout the
        device kernels rely on to split eps / emitting spans without a scan."""
        off = self.row_off
        il = self.arcs["ilabel"]
        is_eps = (il == 0).astype(np.int64)
        csum = np.concatenate([[0], np.cumsum(is_eps)])
        n_eps_row = csum[off[1:]] - csum[off[:-1]]
        if not np.array_equal(n_eps_row, self.niepsilons.astype(np.int64)):
            return False
        # eps count inside the first niepsilons arcs of each row must equal niepsilons
        head_end = off[:-1] + self.niepsilons.astype(np.int64)
        n_eps_head = csum[head_end] - csum[off[:-1]]
        return bool(np.array_equal(n_eps_head, n_eps_row))


def write_fst(path: str, fst: Fst) -> None:
    hdr = np.array(
        [fst.start, fst.final_state, fst.total_states, fst.total_arcs,
         int(fst.niepsilons.sum()), int(fst.noepsilons.sum())], dtype="<i4")
    info = np.empty(fst.total_states, dtype=STATEINFO_DTYPE)
    info["num_arcs"] = fst.num_arcs
    info["niepsilons"] = fst.niepsilons
    info["noepsilons"] = fst.noepsilons
    with open(path, "wb") as f:
        f.write(hdr.tobytes())
        f.write(info.tobytes())
        f.write(np.ascontiguousarray(fst.arcs).tobytes())


def read_fst(path: str) -> Fst:
    with open(path, "rb") as f:
        raw = f.read(24)
        if len(raw) != 24:
            raise IOError(f"{path}: truncated newfst header")
        hdr = np.frombuffer(raw, dtype="<i4")
        start, final_state, n_states, n_arcs = (int(x) for x in hdr[:4])
        raw_info, raw_arcs = f.read(12 * n_states), f.read(16 * n_arcs)
    if len(raw_info) != 12 * n_states or len(raw_arcs) != 16 * n_arcs:
        raise IOError(f"{path}: truncated newfst file")
    info = np.frombuffer(raw_info, dtype=STATEINFO_DTYPE)
    arcs = np.frombuffer(raw_arcs, dtype=ARC_DTYPE)
    if int(info["num_arcs"].sum()) != n_arcs:
        raise IOError(f"{path}: state arc counts do not sum to total_arcs")
    return Fst(start, final_state, arcs.copy(), info["num_arcs"].copy(),
               info["niepsilons"].copy(), info["noepsilons"].copy())


def write_loglikes(path: str, utts) -> None:
    """``utts``: iterable of float32 [T, P] matrices."""
    utts = list(utts)
    with open(path, "wb") as f:
        f.write(np.array([LL_MAGIC, len(utts)], dtype="<i4").tobytes())
        for m in utts:
            m = np.ascontiguousarray(m, dtype="<f4")
            f.write(np.array([m.shape[0], m.shape[1]], dtype="<i4").tobytes())
            f.write(m.tobytes())


def read_loglikes(path: str):
    out = []
    with open(path, "rb") as f:
        magic, n = np.frombuffer(f.read(8), dtype="<i4")
        if int(magic) != LL_MAGIC:
            raise IOError(f"{path}: bad loglikes magic")
        for _ in range(int(n)):
            t, p = (int(x) for x in np.frombuffer(f.read(8), dtype="<i4"))
            out.append(np.frombuffer(f.read(4 * t * p), dtype="<f4").reshape(t, p).copy())
    return out


# --------------------------------------------------------------------------- OpenFst const files
#
# ``ConstFst<StdArc, int>::Read`` (reference ``src/newfst/const-fst.h:46-80,189-221``) followed by
# ``Fst(const ConstFst&)`` (``src/newfst/optimize-fst.h:82-134``): a real Kaldi ``HCLG.fst`` saved as
# an OpenFst "const" FST becomes the flat graph the decoders read — one super-final state is
# appended and every originally-final state gets ``0:0/final_weight -> super-final`` as its FIRST arc.
# Like the reference's reader, no alignment padding is handled (files written with ``--align`` are
# refused by the magic/size checks downstream).

FST_MAGIC = 2125659606
CONST_STATE_DTYPE = np.dtype([("weight", "<f4"), ("pos", "<i4"), ("narcs", "<i4"), ("niepsilons", "<i4"),
                              ("noepsilons", "<i4")])


def _read_string(f) -> str:
    n = int(np.frombuffer(f.read(4), dtype="<i4")[0])
    return f.read(n).decode()


def read_const_fst(path: str) -> Fst:
    with open(path, "rb") as f:
        if int(np.frombuffer(f.read(4), dtype="<i4")[0]) != FST_MAGIC:
            raise IOError(f"{path}: bad FST header")
        fsttype, arctype = _read_string(f), _read_string(f)
        if fsttype != "const":
            raise IOError(f"{path}: FST not of type const")
        if arctype != "standard":
            raise IOError(f"{path}: arc type {arctype!r}, want standard")
        np.frombuffer(f.read(8), dtype="<i4")                 # version, flags
        np.frombuffer(f.read(8), dtype="<u8")                 # properties
        start, n_states, n_arcs = (int(x) for x in np.frombuffer(f.read(24), dtype="<i8"))
        states = np.frombuffer(f.read(CONST_STATE_DTYPE.itemsize * n_states), dtype=CONST_STATE_DTYPE)
        arcs_in = np.frombuffer(f.read(16 * n_arcs), dtype=ARC_DTYPE)
    if states.shape[0] != n_states or arcs_in.shape[0] != n_arcs:
        raise IOError(f"{path}: truncated const FST")
    is_final = ~np.isposinf(states["weight"])                 # Weight::Zero() is +inf
    n_final = int(is_final.sum())
    S = n_states + 1
    num_arcs = np.zeros(S, np.uint32)
    nie = np.zeros(S, np.uint32)
    noe = np.zeros(S, np.uint32)
    num_arcs[:n_states] = states["narcs"] + is_final
    nie[:n_states] = states["niepsilons"] + is_final
    noe[:n_states] = states["noepsilons"] + is_final
    arcs = np.zeros(n_arcs + n_final, ARC_DTYPE)
    shift = np.cumsum(is_final) - is_final                    # final arcs inserted before each state's row
    dst_pos = states["pos"].astype(np.int64) + shift + is_final
    # scatter every state's arcs behind its (optional) final arc
    src_idx = np.arange(n_arcs, dtype=np.int64)
    owner = np.repeat(np.arange(n_states), states["narcs"])
    arcs[dst_pos[owner] + (src_idx - states["pos"][owner])] = arcs_in
    fpos = (states["pos"].astype(np.int64) + shift)[is_final]
    arcs["weight"][fpos] = states["weight"][is_final]
    arcs["nextstate"][fpos] = n_states
    return Fst(start, n_states, arcs, num_arcs, nie, noe)


def write_const_fst(path: str, fst: Fst) -> None:
    """Inverse of :func:`read_const_fst` (test fixture writer): the super-final state and the
    leading final arcs of a flat graph become OpenFst final weights again."""
    S = fst.total_states - 1
    assert fst.final_state == S and int(fst.num_arcs[S]) == 0
    off = fst.row_off
    first = fst.arcs[np.minimum(off[:S], max(fst.total_arcs - 1, 0))]
    is_final = (fst.num_arcs[:S] > 0) & (first["ilabel"] == 0) & (first["olabel"] == 0) & (first["nextstate"] == S)
    assert not (fst.arcs["nextstate"] == S)[np.setdiff1d(np.arange(fst.total_arcs), off[:S][is_final])].any(), \
        "arcs into the super-final state must lead their rows"
    states = np.zeros(S, CONST_STATE_DTYPE)
    states["weight"] = np.where(is_final, first["weight"], np.float32(np.inf))
    states["narcs"] = fst.num_arcs[:S] - is_final
    states["niepsilons"] = fst.niepsilons[:S] - is_final
    states["noepsilons"] = fst.noepsilons[:S] - is_final
    states["pos"] = np.concatenate([[0], np.cumsum(states["narcs"])[:-1]])
    keep = np.ones(fst.total_arcs, bool)
    keep[off[:S][is_final]] = False
    arcs = fst.arcs[keep]
    with open(path, "wb") as f:
        f.write(np.array([FST_MAGIC], "<i4").tobytes())
        for s in ("const", "standard"):
            f.write(np.array([len(s)], "<i4").tobytes() + s.encode())
        f.write(np.array([1, 0], "<i4").tobytes())            # version, flags (no symbol tables, not aligned)
        f.write(np.array([0], "<u8").tobytes())               # properties
        f.write(np.array([fst.start, S, arcs.shape[0]], "<i8").tobytes())
        f.write(states.tobytes())
        f.write(np.ascontiguousarray(arcs).tobytes())


# --------------------------------------------------------------------------- CLG + HMM set

def write_hmm_set(path: str, hmms) -> None:
    """``ClgFst::ReadHmm`` format (reference ``src/my-decoder/clg-fst.h:49-74``): int32 count, then
    that many newfst graphs back to back; HMM ``i`` of the file gets id ``i + 1`` (a CLG arc's
    ilabel)."""
    with open(path, "wb") as f:
        f.write(np.array([len(hmms)], dtype="<i4").tobytes())
        for h in hmms:
            hdr = np.array([h.start, h.final_state, h.total_states, h.total_arcs,
                            int(h.niepsilons.sum()), int(h.noepsilons.sum())], dtype="<i4")
            info = np.empty(h.total_states, dtype=STATEINFO_DTYPE)
            info["num_arcs"], info["niepsilons"], info["noepsilons"] = h.num_arcs, h.niepsilons, h.noepsilons
            f.write(hdr.tobytes())
            f.write(info.tobytes())
            f.write(np.ascontiguousarray(h.arcs).tobytes())


def read_hmm_set(path: str):
    out = []
    with open(path, "rb") as f:
        n = int(np.frombuffer(f.read(4), "<i4")[0])
        for _ in range(n):
            hdr = np.frombuffer(f.read(24), "<i4")
            ns, na = int(hdr[2]), int(hdr[3])
            info = np.frombuffer(f.read(12 * ns), STATEINFO_DTYPE)
            arcs = np.frombuffer(f.read(16 * na), ARC_DTYPE)
            out.append(Fst(int(hdr[0]), int(hdr[1]), arcs.copy(), info["num_arcs"].copy(),
                           info["niepsilons"].copy(), info["noepsilons"].copy()))
    return out


@dataclasses.dataclass
class ClgGraph:
    """What ``materialize_clg`` returns: the static graph over the reference's own two-level state
    ids plus, per arc, the two weights the best-token pre-pass adds one after the other."""
    fst: Fst
    w_clg: np.ndarray   # f32 [A]: weight of the CLG arc an arc out of a CLG state entered through (else 0)
    w_hmm: np.ndarray   # f32 [A]: weight of the HMM arc itself
    from_clg: np.ndarray  # bool [A]: arc leaves a CLG state through an HMM (two-weight arc)
    offset: int


def materialize_clg(clg: Fst, hmms) -> ClgGraph:
    """The graph ``ClgFst`` (reference ``src/my-decoder/clg-fst.h:9-189``) expands on the fly, written
    out as one static graph over the SAME state ids: CLG state ``s`` keeps its id; the copy of HMM
    state ``k`` inside CLG arc ``a`` is ``a + offset * (k + 1)`` with ``offset = total_arcs + 1``
    (``GetState`` / ``MapClgTokenStateId``, clg-fst.h:82-165).  Arcs, in the order the reference's
    decoder enumerates them (``online-clg-decoder-mempool-base.h:122-205``):

    * CLG state, eps arc (ilabel 0): unchanged;
    * CLG state, arc ``a`` with HMM id ``h``: one arc per EMITTING arc ``e`` of state 0 of HMM ``h``
      — ilabel ``e.ilabel`` (a pdf), olabel of the CLG arc, weight ``e.w + clg.w`` (that order),
      destination ``a + offset`` when ``e`` is the self-loop of state 0, else ``a + 2 offset``;
    * HMM copy ``(a, k)``: the arcs of state ``k`` of the HMM, olabels removed (``RmOlalel``):
      emitting arcs stay (self-loop) or go to ``(a, k + 1)`` whatever their target, the eps arc
      (HMM end) goes to the CLG arc's destination.
    Pure Python loops: test infrastructure for small graphs (the library has its own, asrd_graph_read_clg)."""
    off = clg.row_off
    A = clg.total_arcs
    offset = A + 1
    kmax = max([h.total_states for h in hmms] + [1])
    n_ids = offset * (kmax + 1)
    rows = [[] for _ in range(n_ids)]   # (ilabel, olabel, weight, next, w_clg, w_hmm, from_clg)
    f32 = np.float32
    for s in range(clg.total_states):
        for a in range(off[s], off[s + 1]):
            arc = clg.arcs[a]
            il = int(arc["ilabel"])
            if il == 0:
                rows[s].append((0, int(arc["olabel"]), f32(arc["weight"]), int(arc["nextstate"]), f32(0), f32(arc["weight"]), False))
                continue
            h = hmms[il - 1]
            hoff = h.row_off
            for e in range(hoff[0], hoff[1]):
                ea = h.arcs[e]
                if int(ea["ilabel"]) == 0:
                    continue
                dst = a + offset if int(ea["nextstate"]) == 0 else a + 2 * offset
                rows[s].append((int(ea["ilabel"]), int(arc["olabel"]), f32(f32(ea["weight"]) + f32(arc["weight"])), dst,
                                f32(arc["weight"]), f32(ea["weight"]), True))
            for k in range(h.total_states):
                sid = a + offset * (k + 1)
                for e in range(hoff[k], hoff[k + 1]):
                    ea = h.arcs[e]
                    if int(ea["ilabel"]) == 0:
                        rows[sid].append((0, 0, f32(ea["weight"]), int(arc["nextstate"]), f32(0), f32(ea["weight"]), False))
                    else:
                        dst = sid if int(ea["nextstate"]) == k else sid + offset
                        rows[sid].append((int(ea["ilabel"]), 0, f32(ea["weight"]), dst, f32(0), f32(ea["weight"]), False))
    num = np.array([len(r) for r in rows], dtype=np.uint32)
    nie = np.array([sum(1 for x in r if x[0] == 0) for r in rows], dtype=np.uint32)
    noe = np.array([sum(1 for x in r if x[1] == 0) for r in rows], dtype=np.uint32)
    flat = [x for r in rows for x in r]
    arcs = np.zeros(len(flat), dtype=ARC_DTYPE)
    arcs["ilabel"] = [x[0] for x in flat]
    arcs["olabel"] = [x[1] for x in flat]
    arcs["weight"] = [x[2] for x in flat]
    arcs["nextstate"] = [x[3] for x in flat]
    g = Fst(clg.start, clg.final_state, arcs, num, nie, noe)
    return ClgGraph(g, np.array([x[4] for x in flat], np.float32), np.array([x[5] for x in flat], np.float32),
                    np.array([x[6] for x in flat], bool), offset)
