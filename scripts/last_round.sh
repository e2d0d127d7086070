# under gpurun: the closing measurements of round 2 — lattice lines (single calls and the batched
# call), time split of the batched call, then the whole GPU suite on the final tree
set -x
mkdir -p gpurun_out
timeout 150 python bench.py --workload lattice --steps 3 --warmup 2 2>/dev/null | tail -1 > gpurun_out/r02_bench_lattice.json
timeout 100 python bench.py --workload lattice --steps 3 --warmup 2 --prune-tokens 1 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02_bench_lattice_prune.json
timeout 100 python bench.py --workload lattice --steps 3 --warmup 2 --lattice-utts 256 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02_bench_lattice_256.json
timeout 100 python scripts/lattice_batch_dbg.py > gpurun_out/r02y_lattice_batch_dbg.txt 2>&1; tail -2 gpurun_out/r02y_lattice_batch_dbg.txt
for f in lattice lattice_prune lattice_256; do python -c "import json,sys; d=json.load(open('gpurun_out/r02_bench_$f.json')); print('$f', round(d['value'],3), d['batched'])"; done
timeout 330 python -m pytest tests -m gpu -x -q > gpurun_out/r02y_tests_full.log 2>&1; tail -3 gpurun_out/r02y_tests_full.log
