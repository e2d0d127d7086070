#!/bin/bash
# A/B of environment settings on the same box (run under gpurun): scripts/ab_env.sh "VAR=a" "VAR=b" ...
for rep in 1 2 3; do
for v in "$@"; do
  echo -n "$v: "
  env $v timeout 250 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2))"
done
done
