#!/bin/bash
# Run under gpurun: the GPU test suite (every test under its own time limit, the whole suite under
# a hard one so a hung kernel cannot hold the box), then a short bench.  Logs into gpurun_out/.
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout -k 10 ${TEST_LIMIT:-1500} python -m pytest tests -m gpu -q --timeout 400 ${PYTEST_ARGS} > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_tests.log
tail -25 gpurun_out/${TAG}_tests.log
timeout -k 10 600 python bench.py --steps ${STEPS:-4} --warmup 3 ${BENCH_ARGS} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"
tail -c 3000 gpurun_out/${TAG}_bench.json
tail -5 gpurun_out/${TAG}_bench.err
