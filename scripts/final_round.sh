#!/bin/bash
# Run under gpurun: the round's closing measurements.  GPU tests, the ncu artefacts of the frozen
# kernels (profile_ncu.sh), then every bench line DESIGN.md section 6 quotes, into gpurun_out/<tag>_*.
TAG=${1:-r02v}
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_tests.log
bash scripts/profile_ncu.sh $TAG k_stream > gpurun_out/${TAG}_profile.log 2>&1
line() {  # name, bench arguments
  local name=$1; shift
  timeout 700 python bench.py "$@" > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
  echo "$name exit $? $(tail -1 gpurun_out/${TAG}_bench_${name}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['metric'], round(d['value'],1), 'ms/step', round(d['ms_per_step'],2))" 2>/dev/null)"
}
line offline --steps 5 --warmup 3
line peaked --steps 3 --warmup 3 --regime peaked --cpu-sample-utts 32
line deployed --steps 3 --warmup 3 --regime deployed --cpu-sample-utts 32
line streaming --workload streaming --steps 2 --warmup 1
line streaming_prune25 --workload streaming --steps 1 --warmup 1 --prune-tokens 1
line streaming_prune90 --workload streaming --steps 1 --warmup 1 --prune-tokens 1 --prune-interval 90
line biglm --workload biglm --steps 2 --warmup 1
line lattice --workload lattice --steps 2 --warmup 1
line lattice_prune --workload lattice --steps 2 --warmup 1 --prune-tokens 1
line clg --workload clg --steps 3 --warmup 2
