"""Timeline of one end-to-end (host rows) step: issue times of the chunk copies vs completion."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from asr_decoder_b200 import synth
from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, LatticeFasterDecoderConfig
n, T, P = 256, 333, 3000
fst = synth.make_graph(1_000_000, 5.0, P, seed=12345)
g = CudaFst(fst)
BEAM = float(os.environ.get('BEAM', '13.0'))
cfg = LatticeFasterDecoderConfig(beam=BEAM, max_active=7000, min_active=200, lattice_beam=8.0)
host = torch.empty((n, T, P), dtype=torch.float32).pin_memory()
gen = [synth.make_loglikes(T, P, 2.0, seed=100 + i) for i in range(8)]
for i in range(n):
    host[i] = torch.from_numpy(gen[i % 8])
dev = host.cuda()
batch = CudaDecoderBatch(g, cfg, n, max_frames=T + 8, token_capacity=T * 12000)
def arrays(base):
    ptrs = (C.c_void_p * n)(*[base + i * T * P * 4 for i in range(n)])
    return ptrs, (C.c_int32 * n)(*([T] * n)), (C.c_int32 * n)(*([P] * n))
def step(a, on_device):
    batch.InitDecoding(None); batch.AdvanceDecodingRaw(a[0], a[1], a[2], P, on_device, -1, None)
    batch.FinalizeDecoding(None); return batch.GetBestPath(True, None, vectors=False)
ha, da = arrays(host.data_ptr()), arrays(dev.data_ptr())
for _ in range(2): step(da, True); step(ha, False)
for name, a, od in (("resident", da, True), ("host", ha, False)):
    torch.cuda.synchronize(); t0 = time.perf_counter(); step(a, od); torch.cuda.synchronize()
    print(name, "step ms", round(1e3 * (time.perf_counter() - t0), 2), flush=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    batch.InitDecoding(None); t1 = time.perf_counter()
    batch.AdvanceDecodingRaw(a[0], a[1], a[2], P, od, -1, None); t2 = time.perf_counter()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    batch.FinalizeDecoding(None); r = batch.GetBestPath(True, None, vectors=False); t4 = time.perf_counter()
    print("   init %.2f  advance-issue %.2f  advance-sync %.2f  bestpath %.2f ms" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3)), flush=True)
def timed(a, od, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); step(a, od); torch.cuda.synchronize()
        best = min(best, 1e3 * (time.perf_counter() - t0))
    return round(best, 2)
print("beam", BEAM, "resident", timed(da, True), "host", timed(ha, False), flush=True)
for ch in (8, 16, 32, 64):
    os.environ["ASRD_HOST_CHUNK"] = str(ch)
    print("host chunk", ch, timed(ha, False), flush=True)
os.environ["ASRD_HOST_CHUNK"] = "16"
os.environ["ASRD_TRACE"] = "1"
step(ha, False)
