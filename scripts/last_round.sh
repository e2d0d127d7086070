# under gpurun: the closing measurements of round 2 (offline line with the current traffic evidence,
# memcheck of the small all-routes run, time split of the batched lattice call)
set -x
mkdir -p gpurun_out
timeout 200 python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02_bench_offline.json
head -c 300 gpurun_out/r02_bench_offline.json; echo
timeout 200 python scripts/lattice_batch_dbg.py > gpurun_out/r02y_lattice_batch_dbg.txt 2>&1; tail -3 gpurun_out/r02y_lattice_batch_dbg.txt
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/r02y_sanitizer_memcheck.txt 2>&1; tail -3 gpurun_out/r02y_sanitizer_memcheck.txt
