#!/bin/bash
# Runs ON a 1-GPU box: GPU parity suite (with the experimental variants), variant sweep at 256^3, one bench line.
# Usage: tools/gpu_r2_single.sh <tag> [pytest -k expression]
set -u
tag=${1:-r2}
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1
LFK_TEST_EXPERIMENTAL=1 timeout -k 5 ${PYTEST_TIMEOUT:-900} python -m pytest tests -m gpu -q -x ${2:+-k "$2"} > gpurun_out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/${tag}_pytest_gpu.log
if [ "${SKIP_SWEEP:-0}" != "1" ]; then
timeout -k 5 400 python tools/variant_sweep.py --grid 256 --tag ${tag} ${SWEEP_ARGS:-} > gpurun_out/${tag}_sweep.log 2>&1
echo "sweep rc=$?"; cut -c1-420 gpurun_out/${tag}_sweep.log | tail -30
fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
timeout -k 5 400 python bench.py --steps 10 --warmup 3 --ref-grid-large 0 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
fi
