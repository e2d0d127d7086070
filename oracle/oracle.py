"""TEST INFRASTRUCTURE — Python bindings for the parity checkers.

* ``OracleDecoder``: ctypes wrapper over ``oracle/liboracle.so`` (``wfst_oracle.c``, our C
  restatement of the reference decoder).
* ``run_ref``: runs ``oracle/_ref/ref_decode`` (the UNMODIFIED reference, compiled in place
  from /root/reference by ``oracle/Makefile``) on files and parses its JSON lines.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import dataclasses

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "ref_decode")
REF_BIN_BIGLM = os.path.join(HERE, "_ref", "ref_decode_biglm")

MODE_REFERENCE = 0
MODE_CANONICAL = 1


class OrcConfig(C.Structure):
    _fields_ = [("beam", C.c_float), ("max_active", C.c_int32), ("min_active", C.c_int32),
                ("lattice_beam", C.c_float), ("prune_interval", C.c_int32),
                ("beam_delta", C.c_float), ("hash_ratio", C.c_float), ("prune_scale", C.c_float)]


FRAME_STAT_DTYPE = np.dtype([
    ("n_in", "<u4"), ("cur_cutoff", "<f4"), ("abeam", "<f4"), ("next_cutoff", "<f4"),
    ("n_raw", "<u4"), ("n_within", "<u4"), ("best", "<f4"), ("_pad", "<u4"),
    ("ll_calls", "<i8"), ("arcs_expanded", "<i8"), ("arcs_admitted", "<i8"), ("eps_arcs", "<i8")])
LAT_TOK_DTYPE = np.dtype([("frame", "<i4"), ("state", "<i4"), ("tot", "<f4"), ("extra", "<f4")])
LAT_LINK_DTYPE = np.dtype([("src_frame", "<i4"), ("src_state", "<i4"), ("dst_frame", "<i4"),
                           ("dst_state", "<i4"), ("ilabel", "<i4"), ("olabel", "<i4"),
                           ("graph", "<f4"), ("acoustic", "<f4"), ("src_tot", "<f4"), ("dst_tot", "<f4")])

_lib = None


def build(force: bool = False) -> None:
    """Compile the checkers (liboracle.so always; oracle/_ref only where /root/reference exists)."""
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "wfst_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_graph_create.restype = C.c_void_p
        L.orc_graph_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64,
                                       C.c_int32, C.c_int32]
        L.orc_graph_destroy.argtypes = [C.c_void_p]
        L.orc_graph_set_clg.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_decoder_create.restype = C.c_void_p
        L.orc_decoder_create.argtypes = [C.c_void_p, C.POINTER(OrcConfig), C.c_int]
        L.orc_decoder_destroy.argtypes = [C.c_void_p]
        L.orc_lm_create.restype = C.c_void_p
        L.orc_lm_create.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int64]
        L.orc_lm_destroy.argtypes = [C.c_void_p]
        L.orc_decoder_create_biglm.restype = C.c_void_p
        L.orc_decoder_create_biglm.argtypes = [C.c_void_p, C.POINTER(OrcConfig), C.c_int, C.c_void_p, C.c_void_p]
        L.orc_init_decoding.argtypes = [C.c_void_p]
        L.orc_advance_decoding.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.orc_finalize_decoding.argtypes = [C.c_void_p]
        L.orc_num_frames_decoded.argtypes = [C.c_void_p]
        L.orc_num_frames_decoded.restype = C.c_int32
        L.orc_get_best_path.restype = C.c_int32
        L.orc_get_best_path.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int32]
        L.orc_path_to_vector.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.POINTER(C.c_int32), C.c_void_p,
                                         C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_frame_stats.restype = C.c_int32
        L.orc_frame_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_dump_lattice.restype = C.c_int64
        L.orc_dump_lattice.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        _lib = L
    return _lib


def make_config(beam=13.0, max_active=7000, min_active=200, lattice_beam=8.0, prune_interval=25,
                beam_delta=0.5, hash_ratio=2.0, prune_scale=0.1) -> OrcConfig:
    return OrcConfig(beam, max_active, min_active, lattice_beam, prune_interval, beam_delta,
                     hash_ratio, prune_scale)


@dataclasses.dataclass
class OneBest:
    ok: bool
    words: list
    ali: list
    tot: float
    lm: float
    ilabel: np.ndarray = None
    olabel: np.ndarray = None
    graph: np.ndarray = None
    acoustic: np.ndarray = None

    @property
    def tot_bits(self) -> int:
        return int(np.float32(self.tot).view(np.uint32))


def path_to_vector(ilabel, olabel, graph, acoustic):
    """LatticeToVector (reference src/newfst/lattice-functions.cc:179-217) via the C oracle."""
    L = lib()
    n = len(ilabel)
    il = np.ascontiguousarray(ilabel, dtype=np.int32)
    ol = np.ascontiguousarray(olabel, dtype=np.int32)
    g = np.ascontiguousarray(graph, dtype=np.float32)
    a = np.ascontiguousarray(acoustic, dtype=np.float32)
    words = np.zeros(max(n, 1), np.int32)
    ali = np.zeros(max(n, 1), np.int32)
    nw, na = C.c_int32(0), C.c_int32(0)
    tot, lm = C.c_float(0), C.c_float(0)
    L.orc_path_to_vector(il.ctypes.data, ol.ctypes.data, g.ctypes.data, a.ctypes.data, n,
                         words.ctypes.data, C.byref(nw), ali.ctypes.data, C.byref(na),
                         C.byref(tot), C.byref(lm))
    return words[:nw.value].tolist(), ali[:na.value].tolist(), float(tot.value), float(lm.value)


class OracleGraph:
    def __init__(self, fst, clg=None):
        """``clg``: an ``fstio.ClgGraph`` — ``fst`` is then its materialised graph and the decoders
        follow the reference's CLG decoder (see ``orc_graph_set_clg``)."""
        L = lib()
        if clg is not None:
            fst = clg.fst
        self._arcs = np.ascontiguousarray(fst.arcs)
        self._off = np.ascontiguousarray(fst.row_off, dtype=np.int64)
        self._ieps = np.ascontiguousarray(fst.niepsilons, dtype=np.uint32)
        self.h = L.orc_graph_create(self._arcs.ctypes.data, self._off.ctypes.data,
                                    self._ieps.ctypes.data, fst.total_states, fst.total_arcs,
                                    fst.start, fst.final_state)
        if clg is not None:
            wc = np.ascontiguousarray(clg.w_clg, np.float32)
            wh = np.ascontiguousarray(clg.w_hmm, np.float32)
            fc = np.ascontiguousarray(clg.from_clg, np.uint8)
            L.orc_graph_set_clg(self.h, wc.ctypes.data, wh.ctypes.data, fc.ctypes.data)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_graph_destroy(self.h)
            self.h = None


class OracleLm:
    """LM FSA for the biglm oracle (asr_decoder_b200.lm.LmFsa; pass the OLD LM already rescaled by -1)."""

    def __init__(self, lm):
        self._an = np.ascontiguousarray(lm.states["arc_num"], np.int32)
        self._bp = np.ascontiguousarray(lm.states["backoff_prob"], np.float32)
        self._bi = np.ascontiguousarray(lm.states["backoff_id"], np.int32)
        self._arcs = np.ascontiguousarray(lm.arcs)
        self.h = lib().orc_lm_create(lm.bos, lm.eos, len(lm.states), self._an.ctypes.data, self._bp.ctypes.data,
                                     self._bi.ctypes.data, self._arcs.ctypes.data, len(lm.arcs))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_lm_destroy(self.h)
            self.h = None


class OracleDecoder:
    """Mirrors the reference's DecoderItf call sequence (src/my-decoder/decoder-itf.h:10-25).
    With lm1/lm2 it restates OnlineLatticeDecoderMempoolBiglm (…-biglm.h)."""

    def __init__(self, graph: OracleGraph, cfg: OrcConfig = None, mode: int = MODE_CANONICAL,
                 lm1: "OracleLm" = None, lm2: "OracleLm" = None):
        self.graph = graph
        self.cfg = cfg or make_config()
        self._lms = (lm1, lm2)
        if lm1 is not None:
            self.h = lib().orc_decoder_create_biglm(graph.h, C.byref(self.cfg), mode, lm1.h, lm2.h)
        else:
            self.h = lib().orc_decoder_create(graph.h, C.byref(self.cfg), mode)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_decoder_destroy(self.h)
            self.h = None

    def InitDecoding(self):
        lib().orc_init_decoding(self.h)

    def AdvanceDecoding(self, loglikes: np.ndarray, frames_ready: int = None, max_num_frames: int = -1):
        ll = np.ascontiguousarray(loglikes, dtype=np.float32)
        self._keep = ll
        if frames_ready is None:
            frames_ready = ll.shape[0]
        lib().orc_advance_decoding(self.h, ll.ctypes.data, ll.shape[1], frames_ready, max_num_frames)

    def FinalizeDecoding(self):
        lib().orc_finalize_decoding(self.h)

    def NumFramesDecoded(self) -> int:
        return lib().orc_num_frames_decoded(self.h)

    def GetBestPath(self, use_final_probs: bool = True) -> OneBest:
        cap = 4 * (self.NumFramesDecoded() + 16) + 1024
        while True:
            il = np.zeros(cap, np.int32)
            ol = np.zeros(cap, np.int32)
            g = np.zeros(cap, np.float32)
            a = np.zeros(cap, np.float32)
            n = lib().orc_get_best_path(self.h, int(use_final_probs), il.ctypes.data, ol.ctypes.data,
                                        g.ctypes.data, a.ctypes.data, cap)
            if n == -2:
                cap *= 4
                continue
            break
        if n < 0:
            return OneBest(False, [], [], 0.0, 0.0)
        words, ali, tot, lm = path_to_vector(il[:n], ol[:n], g[:n], a[:n])
        return OneBest(True, words, ali, tot, lm, il[:n].copy(), ol[:n].copy(), g[:n].copy(), a[:n].copy())

    def frame_stats(self) -> np.ndarray:
        n = lib().orc_frame_stats(self.h, None, 0)
        out = np.zeros(n, FRAME_STAT_DTYPE)
        assert FRAME_STAT_DTYPE.itemsize == 64
        lib().orc_frame_stats(self.h, out.ctypes.data, n)
        return out

    def counts(self):
        a, b = C.c_int64(0), C.c_int64(0)
        lib().orc_counts(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def dump_lattice(self):
        nt, nl = self.counts()
        toks = np.zeros(nt, LAT_TOK_DTYPE)
        links = np.zeros(nl, LAT_LINK_DTYPE)
        lib().orc_dump_lattice(self.h, toks.ctypes.data, nt, links.ctypes.data, nl)
        return toks, links

    def decode(self, loglikes: np.ndarray, finalize: bool = True) -> OneBest:
        """InitDecoding -> AdvanceDecoding -> FinalizeDecoding -> GetBestPath -> LatticeToVector,
        the sequence of src/kaldi-nnet3bin/kaldi-hclg-my-decoder.cc:97-122."""
        self.InitDecoding()
        self.AdvanceDecoding(loglikes)
        if finalize:
            self.FinalizeDecoding()
        return self.GetBestPath(True)


def run_ref_biglm(graph_path: str, loglikes_path: str, lm1_path: str, lm2_path: str, **cfg):
    """Run the compiled reference biglm decoder (lm1 is scaled by -1 inside, like the reference bin)."""
    if not (os.path.exists(REF_BIN_BIGLM) and os.access(REF_BIN_BIGLM, os.X_OK)):
        raise RuntimeError("oracle/_ref/ref_decode_biglm is not built")
    cmd = [REF_BIN_BIGLM, f"--graph={graph_path}", f"--loglikes={loglikes_path}", f"--lm1={lm1_path}",
           f"--lm2={lm2_path}"]
    for k, v in cfg.items():
        cmd.append(f"--{k.replace('_', '-')}={v}")
    out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    return [json.loads(l) for l in out.decode().splitlines() if l.startswith("{")]


def have_ref_biglm() -> bool:
    return os.path.exists(REF_BIN_BIGLM) and os.access(REF_BIN_BIGLM, os.X_OK)


def have_ref() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def have_ref_clg() -> bool:
    return os.path.exists(REF_BIN + "_clg")


def run_ref(graph_path: str, loglikes_path: str, stats: bool = True, threads: int = 1, lattice=False,
            chunk: int = 0, repeat: int = 1, hmm_path: str = None, **cfg):
    """Run the compiled reference decoder; returns (list of per-utterance dicts, summary dict).
    ``hmm_path``: the CLG decoder (``ref_decode_clg``: ClgFst + OnlineClgLatticeDecoderMempool) on the
    CLG graph ``graph_path`` and that HMM set."""
    if not have_ref() or (hmm_path and not have_ref_clg()):
        raise RuntimeError("oracle/_ref/ref_decode is not built (run `make -C oracle ref` where "
                           "/root/reference exists)")
    cmd = [REF_BIN + ("_clg" if hmm_path else ""), f"--graph={graph_path}", f"--loglikes={loglikes_path}",
           f"--threads={threads}"]
    if hmm_path:
        cmd.append(f"--hmm={hmm_path}")
    if stats:
        cmd.append("--stats")
    if lattice:
        cmd.append("--lattice")
    if chunk:
        cmd.append(f"--chunk={chunk}")
    if repeat != 1:
        cmd.append(f"--repeat={repeat}")
    for k, v in cfg.items():
        cmd.append(f"--{k.replace('_', '-')}={v}")
    out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    utts, summary = [], None
    for line in out.decode().splitlines():
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if d.get("summary"):
            summary = d
        else:
            utts.append(d)
    return utts, summary
