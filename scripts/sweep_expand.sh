#!/bin/bash
run() {
  echo "== $*"
  env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('ms_per_step', round(d['ms_per_step'],2), 'RTFx', round(d['value']), 'e2e ms', round(d.get('e2e',{}).get('ms_per_step',0),2), d.get('e2e',{}).get('matches_resident_run'), 'kernel ms/launch', {k: round(v,4) for k,v in r['kernel_ms_per_launch'].items()}, 'frac', round(r['frac'],4))
    else: print(l.rstrip()[:300])
"
}
run ASRD_SUBBATCH=256
run ASRD_SUBBATCH=128 ASRD_WORKERS=2
run ASRD_SUBBATCH=64 ASRD_WORKERS=4
run ASRD_SUBBATCH=32 ASRD_WORKERS=8
run ASRD_SUBBATCH=64 ASRD_WORKERS=2
