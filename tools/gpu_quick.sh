#!/bin/bash
# Runs ON the GPU box (under gpurun): parity tests, the A/B sweep of kernel variants and the projection-only sweep.
# Usage: tools/gpu_quick.sh <tag> [grids]
set -u
tag=${1:-quick}
grids=${2:-128,256,512,1024}
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 150 --timeout-method=thread > gpurun_out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/${tag}_pytest_gpu.log
timeout -k 10 400 python tools/variant_sweep.py --grid 256 --tag ${tag} > gpurun_out/${tag}_sweep.log 2>&1
echo "sweep rc=$?"; tail -12 gpurun_out/${tag}_sweep.log
timeout -k 10 400 python tools/projection_sweep.py --grids ${grids} --tag ${tag} > gpurun_out/${tag}_projection.log 2>&1
echo "projection rc=$?"; tail -8 gpurun_out/${tag}_projection.log
