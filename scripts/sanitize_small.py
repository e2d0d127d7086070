"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): plain decode, arena prune in
chunks, raw lattice through both extraction kernels, a CLG graph."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asr_decoder_b200 import synth, fstio
from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, LatticeFasterDecoderConfig

fst = synth.make_graph(3000, 5.0, 60, seed=5, p_final=0.2)
lls = [synth.make_loglikes(40, 60, 2.0, seed=i) for i in range(2)]
cfg = LatticeFasterDecoderConfig(beam=12.0, max_active=800, min_active=50, lattice_beam=6.0, prune_interval=10)
g = CudaFst(fst)
a = CudaDecoderBatch(g, cfg, 2, max_frames=48)
want = a.Decode(lls)
lat = [a.GetRawLattice(i) for i in range(2)]
b = CudaDecoderBatch(g, cfg, 2, max_frames=48, prune_tokens=True)
b.InitDecoding()
for k in range(0, 40, 10):
    b.AdvanceDecoding([x[k:k + 10] for x in lls])
b.FinalizeDecoding()
got = b.GetBestPath(True)
assert all(x.words == y.words and x.tot_bits == y.tot_bits for x, y in zip(want, got))
for i in range(2):
    t, l = b.GetRawLattice(i)
    assert t.tobytes() == lat[i][0].tobytes() and l.tobytes() == lat[i][1].tobytes()
for (t, l), w in zip(a.GetRawLatticeBatch(), lat):     # one launch over both streams
    assert t.tobytes() == w[0].tobytes() and l.tobytes() == w[1].tobytes()
os.environ["ASRD_LATTICE_KERNEL"] = "0"
for i in range(2):
    t, l = a.GetRawLattice(i)
    assert t.tobytes() == lat[i][0].tobytes() and l.tobytes() == lat[i][1].tobytes()
for (t, l), w in zip(a.GetRawLatticeBatch(), lat):
    assert t.tobytes() == w[0].tobytes() and l.tobytes() == w[1].tobytes()
clg, hmms = synth.make_clg(300, n_hmms=20, n_pdfs=40, seed=3)
tmp = tempfile.mkdtemp()
fstio.write_fst(tmp + "/c.fst", clg); fstio.write_hmm_set(tmp + "/h.bin", hmms)
gc = CudaFst.ReadClg(tmp + "/c.fst", tmp + "/h.bin")
c = CudaDecoderBatch(gc, LatticeFasterDecoderConfig(beam=12.0, max_active=300, min_active=20, lattice_beam=6.0), 2, max_frames=48)
out = c.Decode([synth.make_loglikes(40, 40, 1.5, seed=9 + i) for i in range(2)])
print("ok", [len(x.words) for x in want], [len(x.ali) for x in out])
# the HBM-map routes: every frame of the on-chip loop redone through the HBM map, then the
# k_expand / k_post kernels on their own, then a biglm decoder
from asr_decoder_b200 import lm as LM
from asr_decoder_b200.decoder import CudaLm
for env in ({"ASRD_DEBUG_FLAGS": "8"}, {"ASRD_STREAM_KERNEL": "0"}):
    os.environ.update(env)
    d = CudaDecoderBatch(g, cfg, 2, max_frames=48)
    r = d.Decode(lls)
    assert all(x.words == y.words and x.tot_bits == y.tot_bits for x, y in zip(want, r)), env
    for k in env:
        del os.environ[k]
lm1, lm2 = LM.make_lm(300, seed=1, order=2, bigram_density=0.05), LM.make_lm(300, seed=2, order=2, bigram_density=0.05)
fb = synth.make_graph(2000, 5.0, 60, seed=6, n_words=300)
e = CudaDecoderBatch(CudaFst(fb), cfg, 2, max_frames=48, old_lm=CudaLm(lm1.Rescale(-1.0)), new_lm=CudaLm(lm2))
rb = e.Decode(lls)
e.GetRawLattice(0)
e.GetRawLatticeBatch()
print("ok hbm routes + biglm", [x.status for x in rb])
