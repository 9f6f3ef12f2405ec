#!/bin/bash
# Runs ON the GPU box: the position-correction variant test, then the A/B timing of that variant at 256^3.
set -u
tag=${1:-r1d}
mkdir -p gpurun_out
timeout -k 5 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variants" --timeout 100 > gpurun_out/${tag}_pytest_variants.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/${tag}_pytest_variants.log
timeout -k 5 120 python tools/variant_sweep.py --grid 256 --tag ${tag} --only defaults,defaults+correct_masks > gpurun_out/${tag}_sweep.log 2>&1
echo "sweep rc=$?"; tail -4 gpurun_out/${tag}_sweep.log | cut -c1-900
