#!/bin/bash
# Other regimes of SURVEY.md §8d next to the default bench line (run under gpurun): the peaked
# log-likelihoods (sigma 3) and the deployed beam (beam 10, lattice-beam 7).
mkdir -p gpurun_out
: > gpurun_out/bench_regimes.jsonl
for args in "--sigma 3.0" "--sigma 2.0 --beam 10 --lattice-beam 7"; do
  timeout 300 python bench.py --steps 3 --warmup 3 --cpu-sample-utts 32 $args 2>/dev/null | tail -1 >> gpurun_out/bench_regimes.jsonl
done
python - <<'PY'
import json
for l in open("gpurun_out/bench_regimes.jsonl"):
    d = json.loads(l)
    print(json.dumps({"workload": d["config"]["workload"], "value": round(d["value"]), "ms_per_step": round(d["ms_per_step"], 2),
                      "e2e": round(d["e2e"]["value"]), "arcs_per_s": d["arcs_expanded_per_s"], "frac": round(d["roofline"]["frac"], 4),
                      "fallback_frames": d["hbm_map_fallback_frames"], "cpu_reference": round(d["cpu_baseline"]["value"], 1)}))
PY
