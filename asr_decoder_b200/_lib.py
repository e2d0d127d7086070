"""ctypes binding of the C ABI declared in ``include/asrd.h``.

The shared library is CUDA-only.  If it is missing, or no sm_100 device is usable, the
product path raises — there is no CPU fallback and nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

# More hardware work queues than the driver's default of 8: the library runs its sub-batch frame
# loops, staging copies and row scatter on a dozen streams, and streams that share a queue
# serialise (see asrd_on_load in csrc/asrd_api.cu).  Read by the driver at context creation.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libasrd_b200.so")

# every symbol include/asrd.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "asrd_strerror", "asrd_abi_version", "asrd_device_count", "asrd_configure_process",
    "asrd_graph_create", "asrd_graph_read", "asrd_graph_read_const", "asrd_graph_read_clg", "asrd_graph_destroy", "asrd_graph_info",
    "asrd_lm_create", "asrd_lm_destroy", "asrd_lm_convert_arpa",
    "asrd_decoder_create", "asrd_decoder_create_biglm", "asrd_decoder_destroy",
    "asrd_init_decoding", "asrd_advance_decoding", "asrd_finalize_decoding",
    "asrd_num_frames_decoded", "asrd_get_best_path", "asrd_path_to_vector",
    "asrd_frame_stats", "asrd_decoder_status", "asrd_synchronize",
    "asrd_host_alloc", "asrd_host_free", "asrd_launch_count", "asrd_last_fallback_frames", "asrd_last_phase_cycles",
    "asrd_last_pruned_tokens", "asrd_last_peak_tokens", "asrd_last_prune_cycles", "asrd_arena_frame_tokens",
    "asrd_get_raw_lattice", "asrd_get_raw_lattice_batch", "asrd_get_counters", "asrd_profile_enable", "asrd_profile_reset", "asrd_profile_get",
]


class AsrdError(RuntimeError):
    def __init__(self, status: int, where: str = ""):
        self.status = status
        msg = lib().asrd_strerror(status).decode() if _lib is not None else str(status)
        super().__init__(f"{where}: asrd status {status} ({msg})")


class asrd_config(C.Structure):
    _fields_ = [("beam", C.c_float), ("max_active", C.c_int32), ("min_active", C.c_int32),
                ("lattice_beam", C.c_float), ("prune_interval", C.c_int32),
                ("beam_delta", C.c_float), ("hash_ratio", C.c_float), ("prune_scale", C.c_float)]


class asrd_device_options(C.Structure):
    _fields_ = [("hash_capacity", C.c_int32), ("token_capacity", C.c_int64),
                ("max_frames", C.c_int32), ("collect_stats", C.c_int32),
                ("lm_pair_capacity", C.c_int32), ("prune_tokens", C.c_int32), ("reserved", C.c_int32 * 2)]


class asrd_frame_stat(C.Structure):
    _fields_ = [("n_in", C.c_uint32), ("cur_cutoff", C.c_float), ("abeam", C.c_float),
                ("next_cutoff", C.c_float), ("n_tokens", C.c_uint32), ("best", C.c_float),
                ("arcs_expanded", C.c_uint32), ("arcs_admitted", C.c_uint32)]


_lib = None


def kernel_source_sha() -> str:
    """sha1 over the CUDA sources and the ABI header: ties a committed ncu capture
    (profiles/stream_traffic.json) to the code it was taken from."""
    import glob
    import hashlib
    h = hashlib.sha1()
    files = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")) + glob.glob(os.path.join(HERE, "csrc", "*.cuh")))
    files.append(os.path.join(os.path.dirname(HERE), "include", "asrd.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def build(force: bool = False) -> str:
    """Compile ``csrc/`` for sm_100a into ``libasrd_b200.so`` (in-tree)."""
    src_dir = os.path.join(HERE, "csrc")
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.check_call(["make", "-s", "-C", src_dir])
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.asrd_strerror.restype = C.c_char_p
    L.asrd_strerror.argtypes = [C.c_int]
    L.asrd_abi_version.restype = C.c_int
    L.asrd_device_count.restype = C.c_int
    L.asrd_graph_create.argtypes = [vp, vp, vp, i32, i64, i32, i32, C.c_int, C.POINTER(vp)]
    L.asrd_graph_read.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.asrd_graph_read_const.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    L.asrd_graph_read_clg.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(vp)]
    L.asrd_graph_destroy.argtypes = [vp]
    L.asrd_graph_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32),
                                  C.POINTER(i64)]
    L.asrd_decoder_create.argtypes = [vp, C.POINTER(asrd_config), C.POINTER(asrd_device_options),
                                      C.POINTER(vp)]
    L.asrd_decoder_destroy.argtypes = [vp]
    L.asrd_lm_create.argtypes = [i32, i32, i32, vp, vp, vp, vp, i64, C.c_int, C.POINTER(vp)]
    L.asrd_lm_destroy.argtypes = [vp]
    L.asrd_lm_convert_arpa.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    L.asrd_decoder_create_biglm.argtypes = [vp, C.POINTER(asrd_config), C.POINTER(asrd_device_options), vp, vp,
                                            C.POINTER(vp)]
    L.asrd_init_decoding.argtypes = [vp, i32, vp]
    L.asrd_advance_decoding.argtypes = [vp, i32, vp, vp, vp, i32, i32, i32, vp]
    L.asrd_finalize_decoding.argtypes = [vp, i32, vp]
    L.asrd_num_frames_decoded.argtypes = [vp]
    L.asrd_num_frames_decoded.restype = i32
    L.asrd_get_best_path.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.asrd_get_raw_lattice.argtypes = [vp, i32, vp, i64, vp, i64, C.POINTER(i64), C.POINTER(i64), vp]
    L.asrd_get_raw_lattice_batch.argtypes = [vp, i32, i32, vp, i64, vp, i64, vp, vp, vp, vp]
    L.asrd_path_to_vector.argtypes = [vp, vp, vp, vp, i32, vp, C.POINTER(i32), vp, C.POINTER(i32),
                                      C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.asrd_frame_stats.argtypes = [vp, vp, i32, vp]
    L.asrd_frame_stats.restype = i32
    L.asrd_decoder_status.argtypes = [vp, vp]
    L.asrd_synchronize.argtypes = [vp]
    L.asrd_host_alloc.argtypes = [C.POINTER(vp), i64]
    L.asrd_host_free.argtypes = [vp]
    L.asrd_launch_count.restype = i64
    L.asrd_last_fallback_frames.restype = i64
    L.asrd_last_pruned_tokens.restype = i64
    L.asrd_arena_frame_tokens.argtypes = [vp, vp, i32, vp]
    L.asrd_arena_frame_tokens.restype = i32
    L.asrd_last_peak_tokens.restype = i64
    L.asrd_get_counters.argtypes = [vp, i32, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), vp]
    L.asrd_profile_enable.argtypes = [C.c_int]
    L.asrd_profile_get.argtypes = [C.POINTER(C.c_double), C.POINTER(i64)]
    _lib = L
    return L


def check(status: int, where: str = "") -> None:
    if status != 0:
        raise AsrdError(status, where)
