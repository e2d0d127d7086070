"""BASELINE.json configs[2] at FULL size: ~50 M states / ~150 M arcs resident in HBM, one-best + raw
lattice of a few streams against the canonical oracle.  One-off run (about 15 GB of host memory,
a minute of graph generation); the default test suite covers the same code at 8 M / 24 M
(tests/test_gpu_fullsize.py::test_config3_large_graph_lattice).  Prints one JSON line."""
import json, os, sys, time, resource
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from asr_decoder_b200 import synth
from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, LatticeFasterDecoderConfig
from oracle import oracle as O

S = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
SIGMA = float(sys.argv[2]) if len(sys.argv) > 2 else 1.2   # flat scores: thousands of tokens per frame survive
T, P, n = 100, 3000, 4
t0 = time.time(); fst = synth.make_graph(S, 3.0, P, seed=77); t_gen = time.time() - t0
lls = [synth.make_loglikes(T, P, SIGMA, seed=10 + i) for i in range(n)]
cfg = LatticeFasterDecoderConfig(beam=13.0, max_active=7000, min_active=200, lattice_beam=8.0)
t0 = time.time(); g = CudaFst(fst); t_up = time.time() - t0
dec = CudaDecoderBatch(g, cfg, n, max_frames=T + 8, collect_stats=True)
dec.Decode(lls)
t0 = time.time(); out = dec.Decode(lls); t_dec = time.time() - t0
og = O.OracleGraph(fst)
same = []
for i in range(n):
    d = O.OracleDecoder(og, O.make_config(cfg.beam, cfg.max_active, cfg.min_active, cfg.lattice_beam), O.MODE_CANONICAL)
    t0 = time.time(); ref = d.decode(lls[i]); t_cpu = time.time() - t0
    bp = out[i]
    t0 = time.time(); toks, links = dec.GetRawLattice(i); t_lat = time.time() - t0
    same.append(bool(bp.words == ref.words and bp.ali == ref.ali and bp.tot_bits == ref.tot_bits
                     and (len(toks), len(links)) == d.counts()))
print(json.dumps({"states": S, "arcs": int(len(fst.arcs)), "graph_device_bytes": int(g.device_bytes()),
                  "graph_gen_s": round(t_gen, 1), "graph_upload_s": round(t_up, 2), "streams": n, "frames": T,
                  "sigma": SIGMA, "gpu_decode_s_all_streams": round(t_dec, 4), "get_raw_lattice_s_one_stream": round(t_lat, 4),
                  "tokens_per_frame_max": int(dec.frame_stats(n - 1)["n_tokens"].max()) if dec.collect_stats else None, "oracle_decode_s_one_stream": round(t_cpu, 3),
                  "one_best_and_lattice_counts_equal_oracle": same,
                  "lattice_states_links_last_stream": [int(len(toks)), int(len(links))],
                  "host_peak_gb": round(resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6, 1)}))
