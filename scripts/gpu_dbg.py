import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from asr_decoder_b200 import synth
from asr_decoder_b200.decoder import *
from oracle import oracle as O
fst = synth.make_graph(10000, 5.0, 200, seed=12345)
ll = synth.make_loglikes(500, 200, 2.0, seed=777)
cfg = LatticeFasterDecoderConfig(beam=13.0, max_active=7000, min_active=200, lattice_beam=8.0)
g = CudaFst(fst)
dec = CudaDecoderBatch(g, cfg, 1, max_frames=512, collect_stats=True)
t=time.time(); bp = dec.Decode([ll])[0]; print('gpu decode s', time.time()-t, 'status', dec.status(0))
t=time.time(); bp = dec.Decode([ll])[0]; print('gpu decode s (2nd)', time.time()-t)
d = O.OracleDecoder(O.OracleGraph(fst), O.make_config(), O.MODE_CANONICAL)
ref = d.decode(ll); rst = d.frame_stats(); st = dec.frame_stats(0)
print('ok', bp.ok, ref.ok, 'tot', bp.tot, ref.tot, 'words eq', bp.words == ref.words, 'ali eq', bp.ali == ref.ali)
print('n arcs path', len(bp.ilabel), len(ref.ilabel))
for k,(a,b) in dict(n=('n_tokens','n_raw'), cur=('cur_cutoff','cur_cutoff'), ab=('abeam','abeam'), nc=('next_cutoff','next_cutoff'), best=('best','best'), nin=('n_in','n_in'), arcs=('arcs_expanded','arcs_expanded')).items():
    x = st[a]; y = rst[b]
    if x.dtype.kind=='f': eq = x.view(np.uint32)==y.view(np.uint32)
    else: eq = x.astype(np.int64)==y.astype(np.int64)
    bad = np.nonzero(~eq)[0]
    print(k, 'all equal' if len(bad)==0 else f'first mismatch at {bad[0]}: gpu {x[bad[0]]} oracle {y[bad[0]]} (n bad {len(bad)})')
