#!/usr/bin/env python
"""bench.py — batched WFST beam-search decode on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic utterances:
InitDecoding -> AdvanceDecoding(all frames) -> FinalizeDecoding -> GetBestPath for every
stream, i.e. the call sequence of the reference's offline bin
(src/kaldi-nnet3bin/kaldi-hclg-my-decoder.cc:97-122).

Workload at N=1 = BASELINE.json configs[1]: synthetic 1M-state / 5M-arc HCLG, 3000 pdfs,
256 utterances x 333 frames (chain, 30 ms per frame), beam 13, max-active 7000, one-best.
With N GPUs every rank holds a graph replica and its own 256-utterance batch (weak
scaling, no collective on the data path).

  value : RTFx with the log-likelihoods already resident in HBM
  e2e   : RTFx through the C ABI with HOST (pinned) log-likelihoods, H2D + D2H inside
  --impl reference : the reference's own CPU decoder (oracle/_ref/ref_decode, compiled from
          /root/reference) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before the CUDA context exists (see _lib.py)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME_SECONDS = 0.03  # chain model, frame subsampling 3


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--states", type=int, default=1_000_000)
    ap.add_argument("--pdfs", type=int, default=3000)
    ap.add_argument("--utts", type=int, default=256)
    ap.add_argument("--frames", type=int, default=333)
    ap.add_argument("--sigma", type=float, default=2.0)
    ap.add_argument("--beam", type=float, default=13.0)
    ap.add_argument("--max-active", type=int, default=7000)
    ap.add_argument("--min-active", type=int, default=200)
    ap.add_argument("--lattice-beam", type=float, default=8.0)
    ap.add_argument("--lattice-utts", type=int, default=16, help="utterances of the lattice workload")
    ap.add_argument("--cpu-sample-utts", type=int, default=0, help="0 = 2 x host cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--hash-capacity", type=int, default=0, help="0 = library default (8 x max-active)")
    ap.add_argument("--workload", default="offline", choices=["offline", "streaming", "biglm", "lattice", "clg"],
                    help="offline = BASELINE.json configs[1] (the headline); streaming = configs[4]: "
                         "--streams concurrent streams fed in --chunk-frames chunks, sharded over the GPUs; "
                         "biglm = configs[3] (on-the-fly LM-difference composition); lattice = configs[2]'s shape "
                         "(raw lattice from the device + the reference's host determinisation)")
    ap.add_argument("--regime", default="busy", choices=["busy", "peaked", "deployed"],
                    help="SURVEY.md section 8d regimes: busy = sigma 2 (default, max-active binding); peaked = sigma 3; "
                         "deployed = beam 10 / lattice-beam 7 (v2-asrbin/conf/decoder.conf)")
    ap.add_argument("--streams", type=int, default=4096, help="streaming: concurrent streams over ALL GPUs")
    ap.add_argument("--chunk-frames", type=int, default=30)
    ap.add_argument("--prune-tokens", type=int, default=0,
                    help="streaming: 1 = arena pruning every --prune-interval frames (PruneActiveTokens, device option "
                         "prune_tokens): memory per stream bounded by the live tokens, at the price of the prune sweeps")
    ap.add_argument("--prune-interval", type=int, default=25, help="LatticeFasterDecoderConfig::prune_interval")
    ap.add_argument("--token-capacity", type=int, default=0, help="token records per stream (0 = sized from the workload)")
    a = ap.parse_args()
    if a.regime == "peaked":
        a.sigma = 3.0
    elif a.regime == "deployed":
        a.beam, a.lattice_beam = 10.0, 7.0
    return a


def workload_name(a):
    if a.workload == "streaming":
        return (f"streaming: synthetic HCLG {a.states} states avg-degree 5, {a.pdfs} pdfs; {a.streams} concurrent streams x "
                f"{a.frames} frames fed in {a.chunk_frames}-frame chunks (30 ms/frame), sigma={a.sigma}; beam={a.beam} "
                f"max-active={a.max_active} min-active={a.min_active}; one-best")
    return (f"synthetic HCLG {a.states} states avg-degree 5, {a.pdfs} pdfs; {a.utts} utts x {a.frames} frames "
            f"(30 ms/frame), sigma={a.sigma}; beam={a.beam} max-active={a.max_active} "
            f"min-active={a.min_active}; one-best")


def make_inputs(a, rank):
    from asr_decoder_b200 import synth
    fst = synth.make_graph(a.states, 5.0, a.pdfs, seed=12345)
    return fst, (lambda i: synth.make_loglikes(a.frames, a.pdfs, a.sigma, seed=100000 * rank + 1000 + i))


# --------------------------------------------------------------------------- reference arm

def cpu_reference_run(a, fst, gen, n_utts, repeats=1, threads=None):
    """Time the reference's CPU decoder on `n_utts` utterances of the workload.
    Returns dict(value=RTFx, seconds, kind, cores, sample)."""
    from asr_decoder_b200 import fstio
    from oracle import oracle as O
    cores = threads or os.cpu_count() or 1
    lls = [gen(i) for i in range(n_utts)]
    audio = n_utts * a.frames * FRAME_SECONDS
    if O.have_ref():
        tmp = tempfile.mkdtemp(prefix="asrd_ref_")
        try:
            gp, lp = os.path.join(tmp, "g.fst"), os.path.join(tmp, "ll.bin")
            fstio.write_fst(gp, fst)
            fstio.write_loglikes(lp, lls)
            secs = []
            answers = None
            for _ in range(repeats):
                answers, summ = O.run_ref(gp, lp, stats=False, threads=cores, beam=a.beam, max_active=a.max_active,
                                          min_active=a.min_active, lattice_beam=a.lattice_beam)
                secs.append(summ["wall_s"])
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        kind = "reference"
    else:
        # the C port in reference token order, one decoder object per thread
        from concurrent.futures import ThreadPoolExecutor
        answers = None
        og = O.OracleGraph(fst)
        cfg = O.make_config(a.beam, a.max_active, a.min_active, a.lattice_beam)
        decs = [O.OracleDecoder(og, cfg, O.MODE_REFERENCE) for _ in range(cores)]

        def work(t):
            for i in range(t, n_utts, cores):
                decs[t].decode(lls[i])
        secs = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            with ThreadPoolExecutor(cores) as ex:
                list(ex.map(work, range(cores)))
            secs.append(time.perf_counter() - t0)
        kind = "port"
    return dict(seconds=secs, value=audio / float(np.mean(secs)), kind=kind, cores=cores, answers=answers,
                sample=f"{n_utts} of the workload's utterances x {a.frames} frames, {cores} threads "
                       f"(one decoder object per thread, shared graph)")


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fst, gen = make_inputs(a, 0)
    cores = os.cpu_count() or 1
    n_utts = a.cpu_sample_utts or max(2 * cores, 8)
    n_utts = min(n_utts, a.utts)
    res = cpu_reference_run(a, fst, gen, n_utts, repeats=a.warmup + a.steps)
    timed = res["seconds"][a.warmup:]
    audio = n_utts * a.frames * FRAME_SECONDS
    value = audio / float(np.mean(timed))
    line = {
        "impl": "reference", "metric": "batched decode RTFx", "value": value, "unit": "x realtime",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * float(np.mean(timed)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a)},
        "cpu_baseline": {"value": value, "unit": "x realtime", "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"]},
        "e2e": {"value": value, "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- B200 arm

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads to the CPUs next to its GPU (NVML's ideal affinity) BEFORE the
    pinned staging memory is allocated: first touch then places the pages on the GPU's NUMA node, so
    eight ranks do not pull their host->device copies across the socket link (round-1 review: 0.963
    end-to-end efficiency at 8 GPUs).  Best effort: containers may restrict the cpuset."""
    if os.environ.get("ASRD_BENCH_NO_BIND"):
        return 0
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w in range(words) for b in range(64) if (mask[w] >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def run_b200_arm(a):
    import torch
    import torch.distributed as dist
    from asr_decoder_b200 import _lib
    from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, LatticeFasterDecoderConfig, LatticeToVector

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    numa_cpus = bind_to_gpu_numa_node(local) if world > 1 else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()

    fst, gen = make_inputs(a, rank)
    n, T, P = a.utts, a.frames, a.pdfs
    # host inputs in pinned memory (e2e source), device copy for the resident-input number
    host = torch.empty((n, T, P), dtype=torch.float32).pin_memory()
    hv = host.numpy()
    for i in range(n):
        hv[i] = gen(i)
    dev = host.to(f"cuda:{local}", non_blocking=False)

    cfg = LatticeFasterDecoderConfig(beam=a.beam, max_active=a.max_active, min_active=a.min_active,
                                     lattice_beam=a.lattice_beam)
    graph = CudaFst(fst, device=local)
    tok_cap = int(min(max(a.frames + 2, 64) * 12000, 1 << 23))
    batch = CudaDecoderBatch(graph, cfg, n, max_frames=T + 8, token_capacity=tok_cap, hash_capacity=a.hash_capacity)
    stream = torch.cuda.current_stream().cuda_stream

    def ptr_arrays(base_ptr, row_stride_elems):
        ptrs = (C.c_void_p * n)(*[base_ptr + 4 * i * T * row_stride_elems for i in range(n)])
        nfr = (C.c_int32 * n)(*([T] * n))
        strides = (C.c_int32 * n)(*([P] * n))
        return ptrs, nfr, strides

    dev_args = ptr_arrays(dev.data_ptr(), P)
    host_args = ptr_arrays(host.data_ptr(), P)

    def step(args, on_device):
        batch.InitDecoding(stream)
        batch.AdvanceDecodingRaw(args[0], args[1], args[2], P, on_device, -1, stream)
        batch.FinalizeDecoding(stream)
        return batch.GetBestPath(True, stream, vectors=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- resident-input number (value)
    for _ in range(a.warmup):
        res_dev = step(dev_args, True)
    bad = [r.status for r in res_dev if not r.ok]
    if bad:
        raise SystemExit(f"decode failed: statuses {sorted(set(bad))}")
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = L.asrd_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        res_dev = step(dev_args, True)
    e1.record()
    barrier()
    launches = L.asrd_launch_count() - l0
    ms_value = max_over_ranks(e0.elapsed_time(e1) / a.steps)
    clocks = sampler.stop() if sampler else None

    # ---- counters + per-kernel device time of one extra (untimed) step
    ae, aa, tk = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    L.asrd_get_counters(batch.handles, n, C.byref(ae), C.byref(aa), C.byref(tk), stream)
    # per-kernel device time: one extra untimed step with the sub-batch pipelining switched off,
    # so that every launch runs alone on the stream its events are recorded on
    L.asrd_profile_reset()
    L.asrd_profile_enable(1)
    saved = os.environ.get("ASRD_SUBBATCH")
    os.environ["ASRD_SUBBATCH"] = "0"
    step(dev_args, True)
    if saved is None:
        del os.environ["ASRD_SUBBATCH"]
    else:
        os.environ["ASRD_SUBBATCH"] = saved
    L.asrd_profile_enable(0)
    kms, kn = (C.c_double * 4)(), (C.c_int64 * 4)()
    L.asrd_profile_get(kms, kn)
    xm, xn = C.c_double(kms[0]), C.c_int64(kn[0])
    fallback_frames = int(L.asrd_last_fallback_frames())
    on_chip = kn[2] > 0   # the on-chip frame loop (k_stream) served this run
    knames = ["k_expand", "k_post", "k_stream"]
    ktot = kms[0] + kms[1] + kms[2] + 1e-12
    if on_chip:
        xm, xn = C.c_double(kms[2]), C.c_int64(kn[2])

    # ---- end-to-end number: host (pinned) log-likelihoods through the C ABI
    e2e = None
    if not a.no_e2e:
        for _ in range(max(1, min(a.warmup, 2))):
            res_host = step(host_args, False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            res_host = step(host_args, False)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ms_e2e = max_over_ranks(1e3 * (t1 - t0) / a.steps)
        same = all(np.array_equal(x.ilabel, y.ilabel) and np.array_equal(x.olabel, y.olabel) and
                   np.array_equal(x.acoustic.view(np.uint32), y.acoustic.view(np.uint32))
                   for x, y in zip(res_dev, res_host))
        d2h = int(sum(16 * len(r.ilabel) for r in res_host) + 8 * n)
        e2e = (ms_e2e, same, d2h)

    audio_all = sum_over_ranks(n * T * FRAME_SECONDS)
    arcs_all = sum_over_ranks(float(ae.value))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    expand_s = xm.value / 1e3
    achieved = 16.0 * ae.value / expand_s / 1e9 if expand_s > 0 else 0.0
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "stream_traffic.json" if on_chip else "expand_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("kernel_src_sha") == _lib.kernel_source_sha():
                traffic = tj.get("dram_bytes_per_launch")
                traffic_note = f"ncu --set full capture {tj.get('ncu_tag')} of these very sources ({tj.get('kernel_src_sha')})"
            else:  # a capture of other sources says nothing about this build
                traffic_note = (f"stale: profiles/{os.path.basename(tpath)} was captured from sources "
                                f"{tj.get('kernel_src_sha')}, this build is {_lib.kernel_source_sha()}")
        except Exception as e:  # noqa: BLE001
            traffic_note = f"unreadable: {e}"
    line = {
        "metric": "batched decode RTFx", "value": audio_all / (ms_value / 1e3), "unit": "x realtime",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "regime": a.regime, "parallelism": f"replica-per-gpu x{world}, streams sharded",
                   "host_affinity": f"rank 0 bound to the {numa_cpus} CPUs next to its GPU" if numa_cpus else "unbound",
                   "l2": "inputs (1.02 GB of log-likelihoods + per-stream maps) exceed the 126 MB L2; no flush needed",
                   "arena_tokens_per_stream": tok_cap},
        "arcs_expanded_per_s": arcs_all / (ms_value / 1e3),
        "arcs_expanded_per_step": arcs_all,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "hbm_map_fallback_frames": fallback_frames,
        "roofline": {"bound": "hbm", "kernel": "k_stream" if on_chip else "k_expand", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     # the same bytes over the driver-visible step time (sub-batches overlapped, every SM busy; the
                     # launch-duration leg above runs one ragged 256-CTA launch at a time on 148 SMs)
                     "achieved_over_step": 16.0 * float(ae.value) / (ms_value / 1e3) / 1e9,
                     "frac_over_step": 16.0 * float(ae.value) / (ms_value / 1e3) / 1e9 / peak,
                     "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                     "algorithmic_bytes_per_arc": 16, "arcs_per_launch": ae.value / max(1, xn.value),
                     "launch_ms": xm.value / max(1, xn.value), "launches_per_step": int(xn.value),
                     "timing": "CUDA events around every launch of one extra step run without sub-batch overlap",
                     "kernel_src_sha": _lib.kernel_source_sha(),
                     "launch": ("k_stream: one CTA per stream, 32 frames per launch" if on_chip else
                                "k_expand: one frame of all streams per launch"),
                     "kernel_ms_per_launch": {k: kms[i] / max(1, kn[i]) for i, k in enumerate(knames)},
                     "kernel_share_of_step": {k: kms[i] / ktot for i, k in enumerate(knames)}},
    }
    if e2e is not None:
        line["e2e"] = {"value": audio_all / (e2e[0] / 1e3), "unit": "x realtime", "ms_per_step": e2e[0],
                       "h2d_bytes_per_step": int(n * T * P * 4), "d2h_bytes_per_step": e2e[2],
                       "matches_resident_run": bool(e2e[1])}
    if not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        k = min(a.cpu_sample_utts or max(2 * cores, 8), n)
        res = cpu_reference_run(a, fst, gen, k, repeats=1)
        line["cpu_baseline"] = {"value": res["value"], "unit": "x realtime", "cores": res["cores"],
                                "kind": res["kind"], "sample": res["sample"]}
        if res.get("answers"):
            # The reference decoded these very utterances: where does the CUDA one-best stand to ITS
            # answer under this one token order (hash ratio 2.0)?  tests/golden/census.json holds the
            # same comparison under seven orders (the reference disagrees with itself on half of them).
            same = cheaper = dearer = 0
            for r, g in zip(res["answers"], res_dev):
                words, ali, tot, _ = LatticeToVector(g.ilabel, g.olabel, g.graph, g.acoustic)
                tb = int(np.float32(tot).view(np.uint32))
                if words == r["words"] and ali == r["ali"] and tb == r["tot_bits"]:
                    same += 1
                elif tot <= r["tot"]:
                    cheaper += 1
                else:
                    dearer += 1
            line["parity_vs_reference"] = {
                "utterances": len(res["answers"]), "identical_one_best": same, "different_cheaper_or_equal_cost": cheaper,
                "different_dearer": dearer, "reference_token_order": "hash_ratio 2.0, one decoder object per thread",
                "note": "the reference's one-best depends on its token visiting order (SURVEY.md Appendix B-2); "
                        "see tests/golden/census.json and DESIGN.md section 5.1"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- streaming (config 5)

def run_streaming_arm(a):
    """BASELINE.json configs[4]: `--streams` concurrent utterance streams (over all GPUs, sharded
    stream_id mod N: strong scaling of one pool), every stream fed `--chunk-frames` frames per
    AdvanceDecoding call like the reference's service loop (kaldi-online-nnet3-my-decoder.cc:10-48).
    A step = InitDecoding, all chunks of all streams, FinalizeDecoding, GetBestPath."""
    import torch
    import torch.distributed as dist
    from asr_decoder_b200 import _lib, synth
    from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, LatticeFasterDecoderConfig

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    numa_cpus = bind_to_gpu_numa_node(local) if world > 1 else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    fst = synth.make_graph(a.states, 5.0, a.pdfs, seed=12345)
    T, P, CH = a.frames, a.pdfs, a.chunk_frames
    from asr_decoder_b200 import sharding
    my_streams = sharding.shard_indices(a.streams, rank, world)      # stream_id mod N
    n = len(my_streams)
    n_distinct = min(n, 256)                                          # distinct utterances; streams reuse them
    host = torch.empty((n_distinct, T, P), dtype=torch.float32).pin_memory()
    hv = host.numpy()
    for i in range(n_distinct):
        hv[i] = synth.make_loglikes(T, P, a.sigma, seed=1000 + i)
    dev = host.to(f"cuda:{local}")
    cfg = LatticeFasterDecoderConfig(beam=a.beam, max_active=a.max_active, min_active=a.min_active,
                                     lattice_beam=a.lattice_beam, prune_interval=a.prune_interval)
    graph = CudaFst(fst, device=local)
    prune = a.prune_tokens != 0
    # unpruned: every token of the utterance stays (~9 k per frame); pruned: the frames since the last
    # prune (one chunk, at most ~15.3 k tokens per frame) plus the thinned-out history
    span = -(-cfg.prune_interval // CH) * CH     # frames between two prunes (a prune runs at the end of a call)
    tok_cap = a.token_capacity or (int((span + 12) * 11000) if prune else int((T + 2) * 9000))
    free0 = torch.cuda.mem_get_info()[0]
    batch = CudaDecoderBatch(graph, cfg, n, max_frames=T + 8, token_capacity=tok_cap, hash_capacity=a.hash_capacity,
                             prune_tokens=prune)
    stream = torch.cuda.current_stream().cuda_stream
    chunks = [(f0, min(CH, T - f0)) for f0 in range(0, T, CH)]

    def args_for(base_ptr, f0, c):
        ptrs = (C.c_void_p * n)(*[base_ptr + 4 * ((i % n_distinct) * T + f0) * P for i in range(n)])
        return ptrs, (C.c_int32 * n)(*([c] * n)), (C.c_int32 * n)(*([P] * n))

    dev_args = [args_for(dev.data_ptr(), f0, c) for f0, c in chunks]
    host_args = [args_for(host.data_ptr(), f0, c) for f0, c in chunks]

    def step(args, on_device, lat=None):
        batch.InitDecoding(stream)
        for k, (ptrs, nfr, strides) in enumerate(args):
            t0 = time.perf_counter()
            batch.AdvanceDecodingRaw(ptrs, nfr, strides, P, on_device, -1, stream)
            if lat is not None:          # per-chunk latency: the caller waits for the chunk like a service would
                torch.cuda.synchronize()
                lat.append(1e3 * (time.perf_counter() - t0))
        batch.FinalizeDecoding(stream)
        return batch.GetBestPath(True, stream, vectors=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=op)
        return float(t.item())

    for _ in range(a.warmup):
        res = step(dev_args, True)
    slab_used = free0 - torch.cuda.mem_get_info()[0]
    bad = [r.status for r in res if not r.ok]
    if bad:
        raise SystemExit(f"decode failed: statuses {sorted(set(bad))}")
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = L.asrd_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        res = step(dev_args, True)
    e1.record()
    barrier()
    launches = L.asrd_launch_count() - l0
    ms_value = reduce(e0.elapsed_time(e1) / a.steps, dist.ReduceOp.MAX if world > 1 else None)
    clocks = sampler.stop() if sampler else None
    ae, aa, tk = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    L.asrd_get_counters(batch.handles, n, C.byref(ae), C.byref(aa), C.byref(tk), stream)
    fallback_frames = int(L.asrd_last_fallback_frames())
    pruned_tokens, peak_tokens = int(L.asrd_last_pruned_tokens()), int(L.asrd_last_peak_tokens())
    # per-chunk latency (device inputs), then end to end from pinned host memory
    lat = []
    step(dev_args, True, lat)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        res_host = step(host_args, False)
    torch.cuda.synchronize()
    ms_e2e = reduce(1e3 * (time.perf_counter() - t0) / a.steps, dist.ReduceOp.MAX if world > 1 else None)
    same = all(np.array_equal(x.ilabel, y.ilabel) for x, y in zip(res, res_host))
    d2h = int(sum(16 * len(r.ilabel) for r in res_host) + 8 * n)
    audio_all = reduce(n * T * FRAME_SECONDS, dist.ReduceOp.SUM if world > 1 else None)
    arcs_all = reduce(float(ae.value), dist.ReduceOp.SUM if world > 1 else None)
    lat_max = reduce(max(lat), dist.ReduceOp.MAX if world > 1 else None)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": "streaming decode RTFx", "value": audio_all / (ms_value / 1e3), "unit": "x realtime",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "regime": a.regime,
                   "parallelism": f"graph replica per GPU, one pool of {a.streams} streams sharded stream_id mod {world}",
                   "streams_per_gpu": n, "chunks_per_utterance": len(chunks),
                   "host_affinity": f"rank 0 bound to the {numa_cpus} CPUs next to its GPU" if numa_cpus else "unbound",
                   "l2": "inputs and per-stream state exceed the 126 MB L2; no flush needed"},
        "arcs_expanded_per_s": arcs_all / (ms_value / 1e3), "gpu_launches": int(launches), "clocks": clocks,
        "hbm_map_fallback_frames": fallback_frames,
        "chunk_latency_ms": {"mean": float(np.mean(lat)), "max_over_ranks": lat_max, "chunk_frames": CH,
                             "audio_ms_per_chunk": 1e3 * CH * FRAME_SECONDS,
                             "note": "one AdvanceDecoding call over all of this GPU's streams, synchronised"},
        "memory_per_stream_bytes": int(slab_used / max(n, 1)),
        "token_arena": {"prune_tokens": bool(prune), "prune_interval": cfg.prune_interval, "capacity_records": tok_cap,
                        "bytes_per_record": 12 if prune else 8, "peak_records_any_stream": peak_tokens,
                        "records_kept_at_the_end_mean": float(tk.value) / max(n, 1),
                        "records_pruned_mean": pruned_tokens / max(n, 1)},
        "memory_note": (f"per stream: token arena ({tok_cap} records) + {T + 8} x {P} log-likelihood history "
                        f"({(T + 8) * ((P + 3) // 4 * 4) * 4} B) + HBM-map fallback structures (~1.8 MB)"
                        + ("; the arena holds the frames since the last prune plus the lattice-beam survivors of the "
                           "history (PruneActiveTokens every prune_interval frames)" if prune else
                           "; the arena keeps every token of the utterance (prune_tokens off)")),
        "e2e": {"value": audio_all / (ms_e2e / 1e3), "unit": "x realtime", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(n * T * P * 4) * world, "d2h_bytes_per_step": d2h * world,
                "matches_resident_run": bool(same)},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- biglm / lattice lines

def run_side_workload(a):
    """One-GPU measurement lines for BASELINE.json configs[3] (biglm) and configs[2] (lattice mode),
    with the reference's CPU implementation timed beside them.  Not the driver's headline."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from asr_decoder_b200 import _lib, fstio, synth, lm as LM
    from asr_decoder_b200.decoder import CudaDecoderBatch, CudaFst, CudaLm, LatticeFasterDecoderConfig, LatticeToVector
    from oracle import oracle as O
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    L = _lib.lib()
    cores = os.cpu_count() or 1
    cfg = LatticeFasterDecoderConfig(beam=a.beam, max_active=a.max_active, min_active=a.min_active,
                                     lattice_beam=a.lattice_beam)
    stream = torch.cuda.current_stream().cuda_stream
    T, P = a.frames, a.pdfs
    if a.workload == "biglm":
        n = a.utts
        fst = synth.make_graph(a.states, 5.0, P, seed=12345)
        nw = 20000
        lm_order = int(os.environ.get("ASRD_BENCH_LM_ORDER", "2"))   # (1: unigram LMs — how much of the step the LM look-ups are)
        lm1 = LM.make_lm(nw, seed=1, order=lm_order, bigram_density=0.002)
        lm2 = LM.make_lm(nw, seed=2, order=lm_order, bigram_density=0.002)
        lls = [synth.make_loglikes(T, P, a.sigma, seed=1000 + i) for i in range(n)]
        graph = CudaFst(fst)
        batch = CudaDecoderBatch(graph, cfg, n, max_frames=T + 8, token_capacity=(T + 2) * 14000,
                                 old_lm=CudaLm(lm1.Rescale(-1.0)), new_lm=CudaLm(lm2))
        dev = [torch.from_numpy(x).cuda() for x in lls]

        def step():
            batch.InitDecoding(stream)
            batch.AdvanceDecoding(dev, stream=stream)
            batch.FinalizeDecoding(stream)
            return batch.GetBestPath(True, stream, vectors=False)
        for _ in range(a.warmup):
            res = step()
        # -7 (no path within lattice-beam of the best final cost) is the reference's own `false`
        # (…-biglm.h:185-191) and not a failure of the search
        bad = [r.status for r in res if not r.ok and r.status != -7]
        if bad:
            raise SystemExit(f"decode failed: statuses {sorted(set(bad))}")
        torch.cuda.synchronize()
        l0 = L.asrd_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            res = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        ae, aa, tk = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        L.asrd_get_counters(batch.handles, n, C.byref(ae), C.byref(aa), C.byref(tk), stream)
        line = {"metric": "biglm decode RTFx", "value": n * T * FRAME_SECONDS / (ms / 1e3), "unit": "x realtime", "n_gpus": 1,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"biglm: synthetic HCLG {a.states} states, {P} pdfs, two bigram LMs over {nw} words "
                                       f"(old LM scaled by -1); {n} utts x {T} frames, sigma={a.sigma}; beam={a.beam} "
                                       f"max-active={a.max_active}; one-best", "kernels": "k_expand<BIGLM> + k_post<BIGLM> (HBM map)"},
                "arcs_expanded_per_s": ae.value / (ms / 1e3), "gpu_launches": int((L.asrd_launch_count() - l0) / a.steps),
                "utterances_with_a_best_path": int(sum(r.ok for r in res))}
        if O.have_ref_biglm() and not a.no_cpu_baseline:
            tmp = tempfile.mkdtemp(prefix="asrd_biglm_")
            try:
                gp = os.path.join(tmp, "g.fst")
                fstio.write_fst(gp, fst)
                l1, l2 = os.path.join(tmp, "lm1"), os.path.join(tmp, "lm2")
                LM.write_lm(l1, lm1)
                LM.write_lm(l2, lm2)
                k = min(cores, n)

                def one(i):
                    lp = os.path.join(tmp, f"ll{i}")
                    fstio.write_loglikes(lp, [lls[i]])
                    return O.run_ref_biglm(gp, lp, l1, l2, beam=a.beam, max_active=a.max_active, min_active=a.min_active,
                                           lattice_beam=a.lattice_beam)[0]
                with ThreadPoolExecutor(k) as ex:   # one reference process per core, one utterance each, concurrently
                    outs = list(ex.map(one, range(k)))
                wall = max(o["seconds"] for o in outs)
                line["cpu_baseline"] = {"value": k * T * FRAME_SECONDS / wall, "unit": "x realtime", "cores": k, "kind": "reference",
                                        "sample": f"{k} utterances, one OnlineLatticeDecoderMempoolBiglm process per core run "
                                                  "concurrently; slowest decode time (graph / LM load excluded)"}
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
        print(json.dumps(line), flush=True)
        return
    if a.workload == "clg":
        # ---- CLG graph + HMM set: what --graph-type=clg decodes.  The CLG graph is
        # a third of config 2's size because every CLG arc expands into a three-state HMM on the device.
        n = a.utts
        n_clg = max(1000, a.states // 3)
        clg, hmms = synth.make_clg(n_clg, n_hmms=2000, n_pdfs=P, avg_deg=5.0, seed=4321, n_words=20000, p_final=0.02, p_eps=0.15)
        tmp = tempfile.mkdtemp(prefix="asrd_clg_")
        try:
            gp, hp = os.path.join(tmp, "clg.fst"), os.path.join(tmp, "hmm.bin")
            fstio.write_fst(gp, clg)
            fstio.write_hmm_set(hp, hmms)
            graph = CudaFst.ReadClg(gp, hp)
            lls = [synth.make_loglikes(T, P, a.sigma, seed=1000 + i) for i in range(min(n, 64))]
            dev = [torch.from_numpy(x).cuda() for x in lls]
            dev = [dev[i % len(dev)] for i in range(n)]
            batch = CudaDecoderBatch(graph, cfg, n, max_frames=T + 8, token_capacity=(T + 2) * 12000)

            def step():
                batch.InitDecoding(stream)
                batch.AdvanceDecoding(dev, stream=stream)
                batch.FinalizeDecoding(stream)
                return batch.GetBestPath(True, stream, vectors=False)
            for _ in range(a.warmup):
                res = step()
            bad = [r.status for r in res if not r.ok and r.status != -7]
            if bad:
                raise SystemExit(f"decode failed: statuses {sorted(set(bad))}")
            torch.cuda.synchronize()
            l0 = L.asrd_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                res = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            ae, aa, tk = C.c_int64(0), C.c_int64(0), C.c_int64(0)
            L.asrd_get_counters(batch.handles, n, C.byref(ae), C.byref(aa), C.byref(tk), stream)
            line = {"metric": "CLG decode RTFx", "value": n * T * FRAME_SECONDS / (ms / 1e3), "unit": "x realtime", "n_gpus": 1,
                    "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": {"workload": f"clg: synthetic CLG {n_clg} states / {clg.total_arcs} arcs + 2000 three-state HMMs over {P} pdfs "
                                           f"(device graph: {graph.device_bytes() >> 20} MB); {n} utts x {T} frames, sigma={a.sigma}; "
                                           f"beam={a.beam} max-active={a.max_active}; one-best",
                               "kernels": "k_stream<SMEM_LL, CLG> (on-chip frame loop)"},
                    "arcs_expanded_per_s": ae.value / (ms / 1e3), "gpu_launches": int((L.asrd_launch_count() - l0) / a.steps),
                    "hbm_map_fallback_frames": int(L.asrd_last_fallback_frames()),
                    "utterances_with_a_best_path": int(sum(r.ok for r in res))}
            if O.have_ref_clg() and not a.no_cpu_baseline:
                k = min(cores, len(lls))
                lp = os.path.join(tmp, "ll.bin")
                fstio.write_loglikes(lp, lls[:k])
                ref, summ = O.run_ref(gp, lp, stats=False, hmm_path=hp, threads=k, beam=a.beam, max_active=a.max_active,
                                      min_active=a.min_active, lattice_beam=a.lattice_beam)
                line["cpu_baseline"] = {"value": k * T * FRAME_SECONDS / summ["wall_s"], "unit": "x realtime", "cores": k,
                                        "kind": "reference",
                                        "sample": f"{k} utterances, OnlineClgLatticeDecoderMempool over ClgFst, one decoder per thread"}
                same = sum(1 for r, g_ in zip(ref, res)
                           if (lambda v: v[0] == r["words"] and int(np.float32(v[2]).view(np.uint32)) == r["tot_bits"])(
                               LatticeToVector(g_.ilabel, g_.olabel, g_.graph, g_.acoustic)))
                line["parity_vs_reference"] = {"utterances": k, "identical_one_best": same}
            print(json.dumps(line), flush=True)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        return
    # ---- lattice mode: config 3's shape (average degree 3), flat scores so that thousands of tokens survive
    binp = os.path.join(ROOT, "oracle", "_ref", "dropin_nbest")
    states = min(a.states, 2_000_000)
    fst = synth.make_graph(states, 3.0, P, seed=777)
    n = min(a.utts, a.lattice_utts)
    Tl = min(T, 100)
    lls = [synth.make_loglikes(Tl, P, 1.2, seed=50 + i) for i in range(n)]
    graph = CudaFst(fst)
    # --prune-tokens 1: PruneActiveTokens every prune_interval frames while decoding, like the reference
    # (its GetRawLattice then copies an already thin lattice; here k_lattice then sweeps a thin arena)
    cfg.prune_interval = a.prune_interval
    batch = CudaDecoderBatch(graph, cfg, n, max_frames=Tl + 8, prune_tokens=bool(a.prune_tokens))
    batch.Decode(lls)
    t_lat, sizes = [], []
    for rep in range(a.warmup + a.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lat = [batch.GetRawLattice(i) for i in range(n)]
        torch.cuda.synchronize()
        if rep >= a.warmup:
            t_lat.append(1e3 * (time.perf_counter() - t0) / n)
        sizes = [(len(x[0]), len(x[1])) for x in lat]
    d2h = int(np.mean([20 * t + 24 * l for t, l in sizes]))
    # the same n lattices through ONE call: one CTA per stream, the host-side ordering on several threads
    t_batch = []
    for rep in range(a.warmup + a.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lat_b = batch.GetRawLatticeBatch()
        torch.cuda.synchronize()
        if rep >= a.warmup:
            t_batch.append(1e3 * (time.perf_counter() - t0) / n)
    batch_same = all(x[0].tobytes() == y[0].tobytes() and x[1].tobytes() == y[1].tobytes() for x, y in zip(lat, lat_b))
    line = {"metric": "raw lattice extraction ms per utterance", "value": float(np.mean(t_lat)), "unit": "ms",
            "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(np.mean(t_lat)) * n,
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"lattice: synthetic HCLG {states} states avg-degree 3, {P} pdfs; {n} utts x {Tl} frames, "
                                   f"sigma=1.2; beam={a.beam} max-active={a.max_active} lattice-beam={a.lattice_beam}",
                       "prune_tokens": bool(a.prune_tokens),
                       "what": "asrd_get_raw_lattice per utterance: k_lattice (link regeneration + lattice-beam prune on the "
                               "device) + D2H of the survivors + host sort"},
            "raw_lattice_states_links_mean": [float(np.mean([s[0] for s in sizes])), float(np.mean([s[1] for s in sizes]))],
            "d2h_bytes_per_utterance": d2h,
            "batched": {"ms_per_utterance": float(np.mean(t_batch)), "utterances_per_call": n,
                        "identical_to_the_single_calls": bool(batch_same),
                        "what": "asrd_get_raw_lattice_batch: all utterances in one launch (one CTA per stream)"}}
    if os.path.exists(binp) and not a.no_cpu_baseline:
        tmp = tempfile.mkdtemp(prefix="asrd_lat_")
        try:
            gp, lp = os.path.join(tmp, "g.fst"), os.path.join(tmp, "l.llb")
            fstio.write_fst(gp, fst)
            fstio.write_loglikes(lp, lls)
            out = {}
            for which in ("cuda", "ref"):
                r = subprocess.run([binp, f"--graph={gp}", f"--loglikes={lp}", f"--decoder={which}", "--nbest=10",
                                    f"--beam={a.beam}", f"--max-active={a.max_active}", f"--min-active={a.min_active}",
                                    f"--lattice-beam={a.lattice_beam}"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                   check=True).stdout.decode()
                out[which] = [json.loads(x) for x in r.splitlines() if x.startswith("{")]
            same = sum([p["words"] for p in c["nbest"]] == [p["words"] for p in r["nbest"]] for c, r in zip(out["cuda"], out["ref"]))
            for which in ("cuda", "ref"):
                o = out[which]
                line[f"reference_post_pass_{which}"] = {
                    "decode_ms": 1e3 * float(np.mean([x["decode_s"] for x in o])),
                    "get_raw_lattice_ms": 1e3 * float(np.mean([x["raw_lattice_s"] for x in o])),
                    "determinize_nbest_ms": 1e3 * float(np.mean([x["determinize_nbest_s"] for x in o])),
                    "raw_states": float(np.mean([x["raw_states"] for x in o])), "det_states": float(np.mean([x["det_states"] for x in o]))}
            line["nbest10_word_sequences_identical"] = f"{same} of {n} utterances"
            line["post_pass_note"] = ("oracle/_ref/dropin_nbest: the drop-in class and the reference decoder behind one DecoderItf*, "
                                      "both through the reference's own DeterminizeLatticeWrapper + NShortestPath (one stream at a time, 1 host core)")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(line), flush=True)


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    elif a.workload == "streaming":
        run_streaming_arm(a)
    elif a.workload in ("biglm", "lattice", "clg"):
        run_side_workload(a)
    else:
        run_b200_arm(a)


if __name__ == "__main__":
    main()
