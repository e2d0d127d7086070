#!/bin/bash
run() {
  echo "== $*"
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('ms_per_step', round(d['ms_per_step'],2), 'RTFx', round(d['value']), {k: round(v,4) for k,v in r['kernel_ms_per_launch'].items()})
    else: print(l.rstrip()[:300])
"
}
run --hash-capacity 65536
run --hash-capacity 32768
run --hash-capacity 131072
