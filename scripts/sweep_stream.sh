#!/bin/bash
# sub-batch / chunk sweep of the on-chip frame loop (run under gpurun)
for cfg in "32 32" "16 32" "64 32" "256 32" "32 16" "32 64" "32 111" "8 32"; do
  set -- $cfg
  echo "SUBBATCH=$1 DEVICE_CHUNK=$2"
  ASRD_SUBBATCH=$1 ASRD_DEVICE_CHUNK=$2 ASRD_WORKERS=32 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   ms/step', round(d['ms_per_step'],2), 'rtfx', round(d['value']))"
done
