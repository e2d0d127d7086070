"""Host-side mirror of the reference decoder interface, over the C ABI.

Names, argument meaning and error behaviour follow the reference:

* ``LatticeFasterDecoderConfig`` — ``src/my-decoder/lattice-faster-decoder-conf.h:8-69``
* ``CudaLatticeDecoder`` — the ``DecoderItf`` surface (``src/my-decoder/decoder-itf.h:10-25``):
  ``InitDecoding / AdvanceDecoding / FinalizeDecoding / NumFramesDecoded / Decode /
  GetBestPath``; constructed from ``(graph, config)`` like
  ``OnlineLatticeDecoderBase(FST*, const LatticeFasterDecoderConfig&)``
  (``online-decoder-base.h:95``).
* ``CudaDecoderBatch`` — the same calls over N decoder objects at once (one launch
  sequence steps every stream): what replaces the reference's one-decoder-per-pthread
  deployment (``src/v2-asrbin/v2-asr-service.cc:95-104``).
* ``LatticeToVector`` — ``src/newfst/lattice-functions.cc:179-217``.

All search work happens in the CUDA library; this module only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import asrd_config, asrd_device_options, asrd_frame_stat, check
from .fstio import Fst, read_fst

LAT_TOKEN_DTYPE = np.dtype([("frame", "<i4"), ("state", "<i4"), ("cost", "<f4"), ("extra", "<f4"),
                            ("is_final", "<i4")])
LAT_LINK_DTYPE = np.dtype([("src", "<i4"), ("dst", "<i4"), ("ilabel", "<i4"), ("olabel", "<i4"),
                           ("graph", "<f4"), ("acoustic", "<f4")])
FRAME_STAT_DTYPE = np.dtype([("n_in", "<u4"), ("cur_cutoff", "<f4"), ("abeam", "<f4"),
                             ("next_cutoff", "<f4"), ("n_tokens", "<u4"), ("best", "<f4"),
                             ("arcs_expanded", "<u4"), ("arcs_admitted", "<u4")])


@dataclasses.dataclass
class LatticeFasterDecoderConfig:
    """Defaults are the reference's (lattice-faster-decoder-conf.h:35-44)."""
    beam: float = 16.0
    max_active: int = 2 ** 31 - 1
    min_active: int = 200
    lattice_beam: float = 10.0
    prune_interval: int = 25
    determinize_lattice: bool = True
    beam_delta: float = 0.5
    hash_ratio: float = 2.0
    prune_scale: float = 0.1

    def Check(self) -> None:
        # lattice-faster-decoder-conf.h:62-67 (the reference asserts)
        assert (self.beam > 0.0 and self.max_active > 1 and self.lattice_beam > 0.0
                and self.prune_interval > 0 and self.beam_delta > 0.0 and self.hash_ratio >= 1.0
                and 0.0 < self.prune_scale < 1.0)

    def to_c(self) -> asrd_config:
        return asrd_config(self.beam, self.max_active, self.min_active, self.lattice_beam,
                           self.prune_interval, self.beam_delta, self.hash_ratio, self.prune_scale)


@dataclasses.dataclass
class BestPath:
    """The linear lattice GetBestPath fills (one arc per back-trace step, path order),
    plus its LatticeToVector view."""
    ok: bool
    status: int
    ilabel: np.ndarray
    olabel: np.ndarray
    graph: np.ndarray
    acoustic: np.ndarray
    words: List[int]
    ali: List[int]
    tot: float
    lm: float

    @property
    def tot_bits(self) -> int:
        return int(np.float32(self.tot).view(np.uint32))


def LatticeToVector(ilabel, olabel, graph, acoustic):
    """words, alignment, tot_score, lm_score of a best path (lattice-functions.cc:179-217)."""
    L = _lib.lib()
    il = np.ascontiguousarray(ilabel, np.int32)
    ol = np.ascontiguousarray(olabel, np.int32)
    g = np.ascontiguousarray(graph, np.float32)
    a = np.ascontiguousarray(acoustic, np.float32)
    n = il.shape[0]
    words = np.zeros(max(n, 1), np.int32)
    ali = np.zeros(max(n, 1), np.int32)
    nw, na = C.c_int32(0), C.c_int32(0)
    tot, lm = C.c_float(0), C.c_float(0)
    check(L.asrd_path_to_vector(il.ctypes.data, ol.ctypes.data, g.ctypes.data, a.ctypes.data, n,
                                words.ctypes.data, C.byref(nw), ali.ctypes.data, C.byref(na),
                                C.byref(tot), C.byref(lm)), "asrd_path_to_vector")
    return words[:nw.value].tolist(), ali[:na.value].tolist(), float(tot.value), float(lm.value)


class CudaFst:
    """Device-resident HCLG (the reference's ``Fst``, optimize-fst.h:53-307, as CSR in HBM)."""

    def __init__(self, fst: Fst, device: int = 0):
        L = _lib.lib()
        self.host = fst
        self.device = device
        arcs = np.ascontiguousarray(fst.arcs)
        na = np.ascontiguousarray(fst.num_arcs, np.uint32)
        ne = np.ascontiguousarray(fst.niepsilons, np.uint32)
        h = C.c_void_p()
        check(L.asrd_graph_create(arcs.ctypes.data, na.ctypes.data, ne.ctypes.data, fst.total_states,
                                  fst.total_arcs, fst.start, fst.final_state, device, C.byref(h)),
              "asrd_graph_create")
        self.h = h
        self.max_ilabel = int(fst.arcs["ilabel"].max()) if fst.total_arcs else 0

    @classmethod
    def ReadFst(cls, path: str, device: int = 0) -> "CudaFst":
        """``Fst::ReadFst`` (optimize-fst.h:208-280)."""
        return cls(read_fst(path), device)

    @classmethod
    def ReadConstFst(cls, path: str, device: int = 0) -> "CudaFst":
        """``ConstFst::Read`` + ``Fst(const ConstFst&)`` (const-fst.h:189-221, optimize-fst.h:82-134):
        an OpenFst const file; the conversion runs inside the library (``asrd_graph_read_const``)."""
        from .fstio import read_const_fst
        self = cls.__new__(cls)
        self.host = None
        self.device = device
        h = C.c_void_p()
        check(_lib.lib().asrd_graph_read_const(path.encode(), device, C.byref(h)), "asrd_graph_read_const")
        self.h = h
        self.max_ilabel = int(read_const_fst(path).arcs["ilabel"].max())
        return self

    @classmethod
    def ReadClg(cls, clg_path: str, hmm_path: str, device: int = 0) -> "CudaFst":
        """``ClgFst::Init(clgfst, hmmfst)`` (my-decoder/clg-fst.h:17-74): the CLG graph and its HMM
        set as one static device graph (``asrd_graph_read_clg``); decoders on it follow the
        reference's CLG decoder (``OnlineClgLatticeDecoderMempool``)."""
        from .fstio import read_hmm_set
        self = cls.__new__(cls)
        self.host = None
        self.device = device
        h = C.c_void_p()
        check(_lib.lib().asrd_graph_read_clg(clg_path.encode(), hmm_path.encode(), device, C.byref(h)),
              "asrd_graph_read_clg")
        self.h = h
        self.max_ilabel = max([int(x.arcs["ilabel"].max()) for x in read_hmm_set(hmm_path) if x.total_arcs] + [0])
        return self

    def device_bytes(self) -> int:
        b = C.c_int64(0)
        check(_lib.lib().asrd_graph_info(self.h, None, None, None, None, C.byref(b)), "asrd_graph_info")
        return b.value

    def close(self):
        if getattr(self, "h", None):
            _lib.lib().asrd_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CudaLm:
    """Device-resident LM FSA (the reference's ``ArpaLm``/``Fsa``, newlm/arpa2fsa.h:217-480).
    Pass the OLD LM already rescaled by -1, like the reference's bin does
    (kaldi-hclg-my-decoder-biglm.cc:55-60)."""

    def __init__(self, lm, device: int = 0):
        L = _lib.lib()
        an = np.ascontiguousarray(lm.states["arc_num"], np.int32)
        bp = np.ascontiguousarray(lm.states["backoff_prob"], np.float32)
        bi = np.ascontiguousarray(lm.states["backoff_id"], np.int32)
        arcs = np.ascontiguousarray(lm.arcs)
        h = C.c_void_p()
        check(L.asrd_lm_create(lm.bos, lm.eos, len(lm.states), an.ctypes.data, bp.ctypes.data, bi.ctypes.data,
                               arcs.ctypes.data, len(arcs), device, C.byref(h)), "asrd_lm_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            _lib.lib().asrd_lm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_ptr_and_keep(x):
    """(device?, pointer, n_frames, stride, keepalive) for a numpy array or a torch tensor."""
    if isinstance(x, np.ndarray):
        a = x if (x.dtype == np.float32 and x.flags.c_contiguous) else np.ascontiguousarray(x, np.float32)
        return False, a.ctypes.data, a.shape[0], a.shape[1], a
    # torch tensor (CUDA or pinned/pageable CPU)
    t = x
    assert t.dim() == 2 and t.dtype.is_floating_point and t.element_size() == 4
    if t.stride(1) != 1:
        t = t.contiguous()
    return bool(t.is_cuda), int(t.data_ptr()), int(t.shape[0]), int(t.stride(0)), t


def get_best_paths(handles, n: int, max_frames_decoded: int, use_final_probs: bool = True, stream: int = 0,
                   vectors: bool = True) -> List[BestPath]:
    """``DecoderItf::GetBestPath`` (inl.h:1071-1200) for ``n`` decoder handles in one batched call."""
    L = _lib.lib()
    cap = 4 * max_frames_decoded + 64
    while True:
        il = np.zeros((n, cap), np.int32)
        ol = np.zeros((n, cap), np.int32)
        gr = np.zeros((n, cap), np.float32)
        ac = np.zeros((n, cap), np.float32)
        na = np.zeros(n, np.int32)
        st = np.zeros(n, np.int32)
        check(L.asrd_get_best_path(handles, n, int(use_final_probs), cap, il.ctypes.data,
                                   ol.ctypes.data, gr.ctypes.data, ac.ctypes.data, na.ctypes.data,
                                   st.ctypes.data, stream), "asrd_get_best_path")
        if (st == -8).any():  # ASRD_ERR_PATH_OVERFLOW
            cap *= 4
            continue
        break
    out = []
    for i in range(n):
        m = int(na[i])
        ok = st[i] == 0 and m > 0
        if ok and vectors:
            words, ali, tot, lm = LatticeToVector(il[i, :m], ol[i, :m], gr[i, :m], ac[i, :m])
        else:
            words, ali, tot, lm = [], [], 0.0, 0.0
        out.append(BestPath(bool(ok), int(st[i]), il[i, :m].copy(), ol[i, :m].copy(), gr[i, :m].copy(),
                            ac[i, :m].copy(), words, ali, tot, lm))
    return out


class CudaDecoderBatch:
    """N decoder objects stepped together.  ``decoders[i]`` is one stream."""

    def __init__(self, graph: CudaFst, config: LatticeFasterDecoderConfig, n: int,
                 max_frames: int = 0, hash_capacity: int = 0, token_capacity: int = 0,
                 collect_stats: bool = False, old_lm: "CudaLm" = None, new_lm: "CudaLm" = None,
                 prune_tokens: bool = False):
        """With ``old_lm``/``new_lm`` the decoders are the biglm variant
        (``OnlineLatticeDecoderMempoolBiglm(&fst, opt, &lm1, &lm2)``).  ``prune_tokens``: drop the
        tokens outside the lattice beam every ``config.prune_interval`` frames (PruneActiveTokens,
        inl.h:438-480); ``token_capacity`` then bounds the live tokens instead of the utterance."""
        config.Check()
        L = _lib.lib()
        self.graph = graph
        self._lms = (old_lm, new_lm)
        self.config = config
        self.n = n
        cfg = config.to_c()
        opts = asrd_device_options(hash_capacity, token_capacity, max_frames, int(collect_stats), 0, int(prune_tokens))
        self.handles = (C.c_void_p * n)()
        self._created = 0
        for i in range(n):
            h = C.c_void_p()
            if old_lm is not None:
                check(L.asrd_decoder_create_biglm(graph.h, C.byref(cfg), C.byref(opts), old_lm.h, new_lm.h,
                                                  C.byref(h)), "asrd_decoder_create_biglm")
            else:
                check(L.asrd_decoder_create(graph.h, C.byref(cfg), C.byref(opts), C.byref(h)),
                      "asrd_decoder_create")
            self.handles[i] = h
            self._created += 1
        self.collect_stats = collect_stats

    def close(self):
        L = _lib.lib()
        for i in range(self._created):
            if self.handles[i]:
                L.asrd_decoder_destroy(self.handles[i])
                self.handles[i] = None
        self._created = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- DecoderItf, batched
    def InitDecoding(self, stream: int = 0) -> None:
        check(_lib.lib().asrd_init_decoding(self.handles, self.n, stream), "asrd_init_decoding")

    def AdvanceDecoding(self, loglikes: Sequence, max_num_frames: int = -1, stream: int = 0) -> None:
        """``loglikes[i]``: rows for the not-yet-decoded frames of stream i — a numpy
        ``[T, P]`` array (host) or a torch tensor (CUDA or CPU); all on the same side."""
        assert len(loglikes) == self.n
        ptrs = (C.c_void_p * self.n)()
        nfr = (C.c_int32 * self.n)()
        strides = (C.c_int32 * self.n)()
        keep = []
        dev = None
        P = None
        for i, x in enumerate(loglikes):
            on_dev, p, t, s, k = _as_ptr_and_keep(x)
            if dev is None:
                dev = on_dev
            assert dev == on_dev, "all log-likelihood buffers must live on the same side"
            cols = int(x.shape[1])
            P = cols if P is None else P
            assert cols == P
            ptrs[i], nfr[i], strides[i] = p, t, s
            keep.append(k)
        if P < self.graph.max_ilabel:
            raise _lib.AsrdError(-1, f"log-likelihood matrix has {P} columns but the graph uses "
                                     f"ilabels up to {self.graph.max_ilabel}")
        self._keep = keep  # host rows must stay alive until the stream is synchronised
        check(_lib.lib().asrd_advance_decoding(self.handles, self.n, ptrs, nfr, strides, P,
                                               max_num_frames, int(dev), stream), "asrd_advance_decoding")

    def AdvanceDecodingRaw(self, ptrs, n_frames, strides, num_indices: int, on_device: bool,
                           max_num_frames: int = -1, stream: int = 0) -> None:
        """Same, with prebuilt ctypes arrays (no per-call marshalling)."""
        check(_lib.lib().asrd_advance_decoding(self.handles, self.n, ptrs, n_frames, strides, num_indices,
                                               max_num_frames, int(on_device), stream), "asrd_advance_decoding")

    def FinalizeDecoding(self, stream: int = 0) -> None:
        check(_lib.lib().asrd_finalize_decoding(self.handles, self.n, stream), "asrd_finalize_decoding")

    def NumFramesDecoded(self, i: int = 0) -> int:
        return _lib.lib().asrd_num_frames_decoded(self.handles[i])

    def Synchronize(self, stream: int = 0) -> None:
        check(_lib.lib().asrd_synchronize(stream), "asrd_synchronize")

    def GetBestPath(self, use_final_probs: bool = True, stream: int = 0, vectors: bool = True) -> List[BestPath]:
        maxf = max(self.NumFramesDecoded(i) for i in range(self.n))
        return get_best_paths(self.handles, self.n, maxf, use_final_probs, stream, vectors)

    def GetRawLattice(self, i: int = 0, use_final_probs: bool = True, stream: int = 0):
        """``DecoderItf::GetRawLattice`` for stream ``i`` (inl.h:868-975), after the lattice-beam
        pruning of ``FinalizeDecoding``: (tokens, links) as structured arrays, tokens sorted by
        (frame, state), ``links['src'/'dst']`` index into tokens.  ``None`` when the reference
        would return false."""
        L = _lib.lib()
        tok_cap, link_cap = 1 << 16, 1 << 17
        while True:
            toks = np.zeros(tok_cap, LAT_TOKEN_DTYPE)
            links = np.zeros(link_cap, LAT_LINK_DTYPE)
            nt, nl = C.c_int64(0), C.c_int64(0)
            rc = L.asrd_get_raw_lattice(self.handles[i], int(use_final_probs), toks.ctypes.data, tok_cap,
                                        links.ctypes.data, link_cap, C.byref(nt), C.byref(nl), stream)
            if rc == -8:  # ASRD_ERR_PATH_OVERFLOW: retry with what the kernel asked for
                tok_cap = max(tok_cap, int(nt.value) + 16)
                link_cap = max(link_cap, int(nl.value) + 16)
                continue
            if rc == -7:
                return None
            check(rc, "asrd_get_raw_lattice")
            return toks[:nt.value].copy(), links[:nl.value].copy()

    def GetRawLatticeBatch(self, use_final_probs: bool = True, stream: int = 0):
        """``GetRawLattice`` of every stream in ONE call (``asrd_get_raw_lattice_batch``: one CTA per
        stream, the host-side ordering on several threads): a list with what ``GetRawLattice(i)``
        returns for each stream, bit for bit."""
        L = _lib.lib()
        # the windows are kept between calls: fresh pages under 256 windows cost more than the call
        tok_cap, link_cap = getattr(self, "_lat_caps", (1 << 15, 1 << 16))
        while True:
            if getattr(self, "_lat_caps", None) != (tok_cap, link_cap):
                self._lat_buf = (np.empty((self.n, tok_cap), LAT_TOKEN_DTYPE), np.empty((self.n, link_cap), LAT_LINK_DTYPE))
                self._lat_caps = (tok_cap, link_cap)
            toks, links = self._lat_buf
            nt, nl = np.zeros(self.n, np.int64), np.zeros(self.n, np.int64)
            st = np.zeros(self.n, np.int32)
            check(L.asrd_get_raw_lattice_batch(self.handles, self.n, int(use_final_probs), toks.ctypes.data, tok_cap,
                                               links.ctypes.data, link_cap, nt.ctypes.data, nl.ctypes.data,
                                               st.ctypes.data, stream), "asrd_get_raw_lattice_batch")
            if (st == -8).any():  # ASRD_ERR_PATH_OVERFLOW: retry with what the kernels asked for
                tok_cap = max(tok_cap, int(nt.max()) + 16)
                link_cap = max(link_cap, int(nl.max()) + 16)
                continue
            out = []
            for i in range(self.n):
                if st[i] == -7:
                    out.append(None)
                    continue
                check(int(st[i]), "asrd_get_raw_lattice_batch")
                out.append((toks[i, :nt[i]].copy(), links[i, :nl[i]].copy()))
            return out

    def arena_frame_tokens(self, i: int = 0, stream: int = 0) -> np.ndarray:
        """Token records the arena holds per frame right now (after the prunes, if any)."""
        L = _lib.lib()
        n = L.asrd_arena_frame_tokens(self.handles[i], None, 0, stream)
        if n < 0:
            raise _lib.AsrdError(n, "asrd_arena_frame_tokens")
        out = np.zeros(max(n, 1), np.uint32)
        if n:
            L.asrd_arena_frame_tokens(self.handles[i], out.ctypes.data, n, stream)
        return out[:n]

    def frame_stats(self, i: int = 0, stream: int = 0) -> np.ndarray:
        L = _lib.lib()
        n = L.asrd_frame_stats(self.handles[i], None, 0, stream)
        if n < 0:
            raise _lib.AsrdError(n, "asrd_frame_stats")
        out = np.zeros(n, FRAME_STAT_DTYPE)
        if n:
            L.asrd_frame_stats(self.handles[i], out.ctypes.data, n, stream)
        return out

    def status(self, i: int = 0, stream: int = 0) -> int:
        return _lib.lib().asrd_decoder_status(self.handles[i], stream)

    def Decode(self, loglikes: Sequence, stream: int = 0) -> List[BestPath]:
        """InitDecoding -> AdvanceDecoding -> FinalizeDecoding -> GetBestPath — the sequence
        every reference caller uses (kaldi-hclg-my-decoder.cc:97-122).  The reference's own
        ``Decode()`` reads one frame past the end (inl.h:615) and is
        not reproduced."""
        self.InitDecoding(stream)
        self.AdvanceDecoding(loglikes, stream=stream)
        self.FinalizeDecoding(stream)
        return self.GetBestPath(True, stream)


class CudaLatticeDecoder:
    """Single-stream drop-in with the ``DecoderItf`` call surface."""

    def __init__(self, graph: CudaFst, config: LatticeFasterDecoderConfig, **device_options):
        self._b = CudaDecoderBatch(graph, config, 1, **device_options)

    def InitDecoding(self) -> None:
        self._b.InitDecoding()

    def AdvanceDecoding(self, loglikes, max_num_frames: int = -1) -> None:
        self._b.AdvanceDecoding([loglikes], max_num_frames)

    def FinalizeDecoding(self) -> None:
        self._b.FinalizeDecoding()

    def NumFramesDecoded(self) -> int:
        return self._b.NumFramesDecoded(0)

    def GetBestPath(self, use_final_probs: bool = True) -> BestPath:
        return self._b.GetBestPath(use_final_probs)[0]

    def Decode(self, loglikes) -> bool:
        bp = self._b.Decode([loglikes])[0]
        self._last = bp
        return bp.ok

    def frame_stats(self) -> np.ndarray:
        return self._b.frame_stats(0)

    def status(self) -> int:
        return self._b.status(0)

    def close(self):
        self._b.close()
