#!/bin/bash
# Run under gpurun.  Produces the ncu launch list (per-launch device time, serialised) and one
# --set full capture of the hot kernels, for profiles/.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
# launch list on the full config-2 batch, a window of launches in the middle of the utterance
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 240 --csv \
    --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
# full-set capture (kernel replay): shorter utterances keep the arenas small
ncu --set full --clock-control none --import-source on -k regex:k_expand -s 60 -c 2 \
    -o gpurun_out/${TAG}_expand -f \
    python bench.py --steps 1 --warmup 1 --frames 100 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_expand.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_boundary -s 60 -c 2 \
    -o gpurun_out/${TAG}_boundary -f \
    python bench.py --steps 1 --warmup 1 --frames 100 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_boundary.log 2>&1
ls -la gpurun_out
