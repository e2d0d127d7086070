#!/bin/bash
# sub-batch / worker / chunk sweep of the on-chip frame loop (run under gpurun)
for cfg in "32 8 32 16" "32 8 16 16" "16 16 32 16" "16 16 16 16" "64 4 32 16" "32 8 32 32" "16 16 16 8" "8 32 16 16"; do
  set -- $cfg
  echo "SUBBATCH=$1 WORKERS=$2 DEVICE_CHUNK=$3 HOST_CHUNK=$4"
  ASRD_SUBBATCH=$1 ASRD_WORKERS=$2 ASRD_DEVICE_CHUNK=$3 ASRD_HOST_CHUNK=$4 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ms/step', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2))"
done
