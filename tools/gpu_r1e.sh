#!/bin/bash
# Runs ON the GPU box: the full GPU parity suite, the A/B timing of the latency-hiding variants at 256^3, and the
# projection-only sweep (BASELINE configs[4]).  Outputs in gpurun_out/.
set -u
tag=${1:-r1e}
mkdir -p gpurun_out
timeout -k 5 70 python -m pytest tests -m gpu -q --timeout 60 > gpurun_out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest_gpu.log
timeout -k 5 40 python tools/variant_sweep.py --grid 256 --tag ${tag} --steps 2 --warmup 3 \
    --only defaults,defaults+correct_prefetch,defaults+g2p_batch,defaults+advect_pair,defaults+all_new > gpurun_out/${tag}_sweep.log 2>&1
echo "sweep rc=$?"; grep -o '"config": "[^"]*", "ms_per_step_wall": [0-9.]*\|"advect_collide": [0-9.]*\|"correct_collide": [0-9.]*\|"g2p": [0-9.]*' gpurun_out/${tag}_sweep.log | tr '\n' ' '; echo
timeout -k 5 30 python tools/projection_sweep.py --grids 128,256,512 --repeats 1 --tag ${tag} > gpurun_out/${tag}_projection.log 2>&1
echo "projection rc=$?"; cut -c1-400 gpurun_out/${tag}_projection.log | tail -4
timeout -k 5 40 python tools/projection_sweep.py --grids 1024 --repeats 1 --tag ${tag}_1024 > gpurun_out/${tag}_projection_1024.log 2>&1
echo "projection 1024 rc=$?"; cut -c1-400 gpurun_out/${tag}_projection_1024.log | tail -2
