import sys, time, os, numpy as np
sys.path.insert(0, '/root/repo')
from asr_decoder_b200 import synth, _lib
from asr_decoder_b200.decoder import *
from oracle import oracle as O
import ctypes as C
S = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
beam = float(sys.argv[4]) if len(sys.argv) > 4 else 13.0
maxa = int(sys.argv[5]) if len(sys.argv) > 5 else 7000
P = 3000
fst = synth.make_graph(S, 5.0, P, seed=12345)
SIG = os.environ.get('SIGMA')
lls = [synth.make_loglikes(T, P, float(SIG) if SIG else (2.0 if i % 2 == 0 else 3.0), seed=4000 + i) for i in range(n)]
cfg = LatticeFasterDecoderConfig(beam=beam, max_active=maxa, min_active=200, lattice_beam=8.0)
g = CudaFst(fst)
dec = CudaDecoderBatch(g, cfg, n, max_frames=T + 8, token_capacity=T * 30000, collect_stats=True)
print('start', flush=True)
t = time.time(); out = dec.Decode(lls); print('gpu decode s', time.time() - t, 'status', [dec.status(i) for i in range(n)], flush=True)
L = _lib.lib()
ae, aa, tk = C.c_int64(0), C.c_int64(0), C.c_int64(0)
L.asrd_get_counters(dec.handles, n, C.byref(ae), C.byref(aa), C.byref(tk), None)
print('arcs', ae.value, aa.value, 'tokens', tk.value, 'fallback frames', L.asrd_last_fallback_frames(), 'of', n * T, flush=True)
ph = (C.c_int64 * 6)(); L.asrd_last_phase_cycles(ph)
totc = sum(ph) or 1
print('phase share: prologue %.3f expansion %.3f closure %.3f write-out %.3f next-cutoff %.3f fallback %.3f | us/frame/stream %.1f' % (*[x / totc for x in ph], totc / 1.965e3 / (n * T)), flush=True)
print('raw phase', list(ph), flush=True)
if os.environ.get('NO_ORACLE'): sys.exit(0)
og = O.OracleGraph(fst)
for i in range(min(n, 2)):
    d = O.OracleDecoder(og, O.make_config(beam=beam, max_active=maxa), O.MODE_CANONICAL)
    ref = d.decode(lls[i]); rst = d.frame_stats(); st = dec.frame_stats(i); bp = out[i]
    print(i, 'ok', bp.ok, ref.ok, 'tot', bp.tot, ref.tot, 'words eq', bp.words == ref.words, 'ali eq', bp.ali == ref.ali)
    for k, (a, b) in dict(n=('n_tokens', 'n_raw'), cur=('cur_cutoff', 'cur_cutoff'), nc=('next_cutoff', 'next_cutoff'), best=('best', 'best'), arcs=('arcs_expanded', 'arcs_expanded')).items():
        x = st[a]; y = rst[b]
        eq = x.view(np.uint32) == y.view(np.uint32) if x.dtype.kind == 'f' else x.astype(np.int64) == y.astype(np.int64)
        bad = np.nonzero(~eq)[0]
        print('  ', k, 'all equal' if len(bad) == 0 else f'first mismatch at {bad[0]}: gpu {x[bad[0]]} oracle {y[bad[0]]} (n bad {len(bad)})', 'max', x.max())
