import sys, numpy as np
sys.path.insert(0, '/root/repo')
from asr_decoder_b200 import synth, lm as LM
from asr_decoder_b200.decoder import *
from oracle import oracle as O
order = int(sys.argv[1]) if len(sys.argv) > 1 else 1
fst = synth.make_graph(3000, 5.0, 90, seed=41, n_words=60, eps_span=250)
lm1 = LM.make_lm(60, seed=7, order=order).Rescale(-1.0)
lm2 = LM.make_lm(60, seed=8, order=order, bigram_density=0.2)
lls = [synth.make_loglikes(t, 90, s, seed=300 + t) for t, s in ((70, 2.0), (33, 2.5), (1, 2.0), (90, 1.8))]
cfg = LatticeFasterDecoderConfig(beam=12.0, max_active=2500, min_active=150, lattice_beam=7.0)
g = CudaFst(fst)
dec = CudaDecoderBatch(g, cfg, len(lls), max_frames=96, collect_stats=True, old_lm=CudaLm(lm1), new_lm=CudaLm(lm2))
out = dec.Decode(lls)
og, o1, o2 = O.OracleGraph(fst), O.OracleLm(lm1), O.OracleLm(lm2)
for i, ll in enumerate(lls):
    d = O.OracleDecoder(og, O.make_config(cfg.beam, cfg.max_active, cfg.min_active, cfg.lattice_beam), O.MODE_CANONICAL, o1, o2)
    ref = d.decode(ll); bp = out[i]; st, rst = dec.frame_stats(i), d.frame_stats()
    print('utt', i, 'status', dec.status(i), 'tot', bp.tot, ref.tot, 'words', bp.words == ref.words, 'ali', bp.ali == ref.ali, 'npath', len(bp.ilabel), len(ref.ilabel))
    print('   n_tokens eq', np.array_equal(st['n_tokens'], rst['n_raw']), 'nc eq', np.array_equal(st['next_cutoff'].view(np.uint32), rst['next_cutoff'].view(np.uint32)), 'best eq', np.array_equal(st['best'].view(np.uint32), rst['best'].view(np.uint32)))
    if len(bp.ilabel) == len(ref.ilabel):
        bad = np.nonzero((bp.ilabel != ref.ilabel) | (bp.olabel != ref.olabel) | (bp.graph.view(np.uint32) != ref.graph.view(np.uint32)) | (bp.acoustic.view(np.uint32) != ref.acoustic.view(np.uint32)))[0]
        print('   bad path positions', bad[:10], 'of', len(bp.ilabel))
        for k in bad[:4]:
            print('      pos', k, 'gpu', bp.ilabel[k], bp.olabel[k], bp.graph[k], bp.acoustic[k], 'ref', ref.ilabel[k], ref.olabel[k], ref.graph[k], ref.acoustic[k])
    else:
        bad = np.nonzero(st['n_tokens'] != rst['n_raw'])[0]; print('  first frame mismatch', bad[:5])
