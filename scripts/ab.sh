#!/bin/bash
# A/B of two prebuilt library variants (asr_decoder_b200/lib_U1.so.variant, lib_U2.so.variant) on
# the same box (run under gpurun); extra bench.py arguments are passed through.
for rep in 1 2 3; do
for v in U1 U2; do
  cp asr_decoder_b200/lib_$v.so.variant asr_decoder_b200/libasrd_b200.so
  echo -n "$v: "
  timeout 250 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), d['hbm_map_fallback_frames'])"
done
done
