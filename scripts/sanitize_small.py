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
os.environ["ASRD_LATTICE_KERNEL"] = "0"
for i in range(2):
    t, l = a.GetRawLattice(i)
    assert t.tobytes() == lat[i][0].tobytes() and l.tobytes() == lat[i][1].tobytes()
clg, hmms = synth.make_clg(300, n_hmms=20, n_pdfs=40, seed=3)
tmp = tempfile.mkdtemp()
fstio.write_fst(tmp + "/c.fst", clg); fstio.write_hmm_set(tmp + "/h.bin", hmms)
gc = CudaFst.ReadClg(tmp + "/c.fst", tmp + "/h.bin")
c = CudaDecoderBatch(gc, LatticeFasterDecoderConfig(beam=12.0, max_active=300, min_active=20, lattice_beam=6.0), 2, max_frames=48)
out = c.Decode([synth.make_loglikes(40, 40, 1.5, seed=9 + i) for i in range(2)])
print("ok", [len(x.words) for x in want], [len(x.ali) for x in out])
