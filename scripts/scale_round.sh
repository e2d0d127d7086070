#!/bin/bash
# Run under `gpurun --gpus N`: the offline (weak scaling) and streaming (strong scaling of one
# 4096-stream pool) workloads on N GPUs, one rank per GPU, lines into gpurun_out/.
N=${1:-2}
TAG=${2:-r02}
mkdir -p gpurun_out
run() {  # name, bench arguments
  local name=$1; shift
  if [ "$N" = 1 ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_${name}_${N}gpu.json 2> gpurun_out/${TAG}_${name}_${N}gpu.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N "$@" > gpurun_out/${TAG}_${name}_${N}gpu.json 2> gpurun_out/${TAG}_${name}_${N}gpu.err
  fi
  echo "$name exit $?"
  tail -1 gpurun_out/${TAG}_${name}_${N}gpu.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$name', 'n_gpus', d['n_gpus'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 1), 'e2e', round(d['e2e']['value']), d['config'].get('host_affinity', ''))"
}
run offline --steps 4 --warmup 3 --no-cpu-baseline
run streaming --workload streaming --steps 2 --warmup 1
