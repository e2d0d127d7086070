#!/bin/bash
# Run under gpurun.  Produces the ncu launch list (per-launch device time, serialised) and one
# --set full capture of each hot kernel, for profiles/.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
shift
KERNELS=${*:-k_expand k_closure k_finalize k_cutoff}
# launch list on the full config-2 batch, a window of launches in the middle of the utterance
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 240 --csv \
    --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
# full-set captures (kernel replay): shorter utterances keep the arenas small
for K in $KERNELS; do
ncu --set full --clock-control none --import-source on -k regex:$K -s 60 -c 1 \
    -o gpurun_out/${TAG}_$K -f \
    python bench.py --steps 1 --warmup 1 --frames 100 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_$K.log 2>&1
done
ls -la gpurun_out
