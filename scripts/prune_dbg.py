"""Arena prune diagnostics on one config-2 stream: tokens kept per frame after every chunk's prune,
time per AdvanceDecoding call with and without pruning."""
import sys, time, os, numpy as np
sys.path.insert(0, '/root/repo')
from asr_decoder_b200 import synth, _lib
from asr_decoder_b200.decoder import *
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 333
n = int(sys.argv[3]) if len(sys.argv) > 3 else 148
CH = 30
P = 3000
fst = synth.make_graph(S, 5.0, P, seed=12345)
lls = [synth.make_loglikes(T, P, 2.0, seed=1000 + i) for i in range(min(n, 16))]
lls = [lls[i % len(lls)] for i in range(n)]
cfg = LatticeFasterDecoderConfig(beam=13.0, max_active=7000, min_active=200, lattice_beam=8.0)
g = CudaFst(fst)
L = _lib.lib()
for prune in (False, True):
    dec = CudaDecoderBatch(g, cfg, n, max_frames=T + 8, token_capacity=T * 12000, prune_tokens=prune)
    for rep in range(2):
        dec.InitDecoding(); L.asrd_synchronize(None)
        times = []; ktimes = []
        import ctypes as C
        L.asrd_profile_enable(1 if os.environ.get('PROFILE') else 0)
        for k in range(0, T, CH):
            L.asrd_profile_reset()
            t0 = time.perf_counter()
            dec.AdvanceDecoding([ll[k:k + CH] for ll in lls]); L.asrd_synchronize(None)
            times.append(1e3 * (time.perf_counter() - t0))
            kms, kn = (C.c_double * 4)(), (C.c_int64 * 4)()
            L.asrd_profile_get(kms, kn)
            ktimes.append((round(kms[2], 1), round(kms[3], 1), kn[2], kn[3]))
            if prune and rep == 1 and k in (0, 30, 60, 150, 300) and not os.environ.get('PROFILE'):
                a = dec.arena_frame_tokens(0)
                print('after frame', k + CH, 'kept per frame (last 45):', a[-45:].tolist(), 'older mean', float(a[:-45].mean()) if len(a) > 45 else None, flush=True)
        if os.environ.get('PROFILE'): print('per chunk (k_stream ms, prune ms, launches):', ktimes, flush=True)
    if prune:
        import ctypes as C
        ae, aa, tk = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        L.asrd_get_counters(dec.handles, n, C.byref(ae), C.byref(aa), C.byref(tk), None)
        pc = (C.c_int64 * 8)(); L.asrd_last_prune_cycles(pc)
        names = ['map', 'emit-links', 'eps-rounds', 'to-front', 'close-up']
        print('prune us per stream per utterance:', {k: round(pc[i] / 1.965e3 / n, 1) for i, k in enumerate(names)},
              'frames swept per stream', pc[6] / n, 'eps rounds', pc[7] / n, flush=True)
    print('prune', prune, 'ms per chunk call (%d streams):' % n, [round(x, 1) for x in times], 'total', round(sum(times), 1), flush=True)
    dec.close()
