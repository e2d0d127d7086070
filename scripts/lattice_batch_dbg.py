"""Where the time of a batched GetRawLattice goes: the C call alone (buffers allocated once) against
the Python method (fresh buffers + per-stream copies), for 16 and 256 utterances."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asr_decoder_b200 import _lib, synth
from asr_decoder_b200.decoder import (LAT_LINK_DTYPE, LAT_TOKEN_DTYPE, CudaDecoderBatch, CudaFst,
                                      LatticeFasterDecoderConfig)

P, T = 3000, 100
fst = synth.make_graph(1_000_000, 3.0, P, seed=777)
g = CudaFst(fst)
cfg = LatticeFasterDecoderConfig(beam=13.0, max_active=7000, min_active=200, lattice_beam=8.0)
L = _lib.lib()
for n in (16, 256):
    lls = [synth.make_loglikes(T, P, 1.2, seed=50 + i) for i in range(n)]
    b = CudaDecoderBatch(g, cfg, n, max_frames=T + 8)
    b.Decode(lls)
    tc, lc = 1 << 15, 1 << 16
    toks, links = np.empty((n, tc), LAT_TOKEN_DTYPE), np.empty((n, lc), LAT_LINK_DTYPE)
    nt, nl, st = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int32)
    out = {"utterances": n}
    for name, fn in (("c_call_reused_buffers_ms", lambda: L.asrd_get_raw_lattice_batch(
                          b.handles, n, 1, toks.ctypes.data, tc, links.ctypes.data, lc, nt.ctypes.data,
                          nl.ctypes.data, st.ctypes.data, 0)),
                     ("python_method_ms", lambda: b.GetRawLatticeBatch())):
        ts = []
        for rep in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0))
        out[name] = [round(x, 2) for x in ts]
    L.asrd_profile_enable(1)
    L.asrd_profile_reset()
    L.asrd_get_raw_lattice_batch(b.handles, n, 1, toks.ctypes.data, tc, links.ctypes.data, lc, nt.ctypes.data,
                                 nl.ctypes.data, st.ctypes.data, 0)
    L.asrd_profile_enable(0)
    out["tokens_links_mean"] = [float(nt.mean()), float(nl.mean())]
    print(json.dumps(out), flush=True)
