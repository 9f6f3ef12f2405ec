#!/bin/bash
# Runs ON the GPU box (under gpurun): GPU parity tests, the default bench line (+ reference arm), the A/B sweep of
# kernel variants and the ncu launch list of the bench command.  Everything lands in gpurun_out/.
# Usage: tools/gpu_r1c.sh <tag>
set -u
tag=${1:-r1c}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout -k 10 420 python -m pytest tests -m gpu -q --timeout 150 --timeout-method=thread > gpurun_out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest_gpu.log
timeout -k 10 300 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json
timeout -k 10 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
echo "bench ref rc=$?"
timeout -k 10 240 python tools/variant_sweep.py --grid 256 --tag ${tag} > gpurun_out/${tag}_sweep.log 2>&1
echo "sweep rc=$?"; tail -9 gpurun_out/${tag}_sweep.log
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/${tag}_launches_256.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline \
    > gpurun_out/${tag}_launches.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out
