#!/bin/bash
run() {
  echo "== $*"
  env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('ms_per_step', round(d['ms_per_step'],2), 'RTFx', round(d['value']), 'e2e ms', round(d.get('e2e',{}).get('ms_per_step',0),2), d.get('e2e',{}).get('matches_resident_run'))
    else: print(l.rstrip()[:300])
"
}
run ASRD_HOST_CHUNK=16
run ASRD_HOST_CHUNK=64
run ASRD_HOST_CHUNK=333
