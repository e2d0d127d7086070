#!/bin/bash
# A/B of prebuilt library variants (asr_decoder_b200/lib_<name>.so.variant) alternated on the same
# box (run under gpurun): scripts/ab.sh "<name> <name> ..." [bench.py arguments].  Box-to-box
# spread is +-1.5 ms, so variants are only ever compared inside one call.
NAMES=${1:-"A B"}
shift
cp asr_decoder_b200/libasrd_b200.so /tmp/libasrd_keep.so
for rep in 1 2 3; do
for v in $NAMES; do
  cp asr_decoder_b200/lib_$v.so.variant asr_decoder_b200/libasrd_b200.so
  echo -n "$v: "
  timeout 250 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), d['hbm_map_fallback_frames'], round(d['roofline']['launch_ms'],3))"
done
done
cp /tmp/libasrd_keep.so asr_decoder_b200/libasrd_b200.so
