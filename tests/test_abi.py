"""The C-ABI shared library loads without a GPU and exports every symbol include/asrd.h declares
(no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "asrd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asrd_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from asr_decoder_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _declared()


def test_library_exports_every_declared_symbol():
    from asr_decoder_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    L = C.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), name
    L.asrd_abi_version.restype = C.c_int
    assert L.asrd_abi_version() == 1
    L.asrd_strerror.restype = C.c_char_p
    assert b"overflow" in L.asrd_strerror(-4)


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU the product path fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from asr_decoder_b200 import _lib, synth
    from asr_decoder_b200.decoder import CudaFst
    with pytest.raises(_lib.AsrdError) as e:
        CudaFst(synth.make_tiny_graph())
    assert e.value.status == -2  # ASRD_ERR_CUDA


def test_path_to_vector_host_helper():
    """LatticeToVector (lattice-functions.cc:179-217): float sums in path order."""
    from asr_decoder_b200.decoder import LatticeToVector
    il = np.array([0, 5, 0, 7, 7], np.int32)
    ol = np.array([0, 0, 9, 0, 3], np.int32)
    g = np.array([0.0, 1.25, 0.5, 2.0, 0.125], np.float32)
    a = np.array([0.0, 3.5, 0.0, 4.75, 1.0], np.float32)
    words, ali, tot, lm = LatticeToVector(il, ol, g, a)
    assert words == [9, 3] and ali == [5, 7, 7]
    t = np.float32(0)
    m = np.float32(0)
    for x, y in zip(g, a):
        m = np.float32(m + x)
        t = np.float32(t + np.float32(x + y))
    assert np.float32(tot) == t and np.float32(lm) == m


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "asr_decoder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "wfst_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
