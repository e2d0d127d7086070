#!/bin/bash
# Run under gpurun.  Produces the ncu launch list (per-launch device time, serialised) and one
# --set full capture of each hot kernel, for profiles/.  Sub-batch overlap is switched off so
# that every launch covers all 256 streams (what bench.py's roofline leg times).
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
shift
KERNELS=${*:-k_stream}
export ASRD_SUBBATCH=0
python -c "from asr_decoder_b200 import _lib; print(_lib.kernel_source_sha())" > gpurun_out/${TAG}_src_sha.txt
# launch list of one full config-2 step (k_begin_advance + one k_stream per 32-frame chunk, back-trace)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_launches.log 2>&1
# full-set capture (kernel replay) of a mid-utterance chunk: the 5th launch of the kernel
for K in $KERNELS; do
ncu --set full --clock-control none --import-source on -k regex:$K -s 5 -c 1 \
    -o gpurun_out/${TAG}_$K -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_$K.log 2>&1
done
ls -la gpurun_out
