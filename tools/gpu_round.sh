#!/bin/bash
# Runs ON the GPU box (under gpurun): GPU parity tests, the default bench line, the ncu launch list of the bench
# command and one `--set full` capture of the particle kernels of a warm step.  Everything lands in gpurun_out/.
# Usage: tools/gpu_round.sh <tag> [skip-tests]
set -u
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_smi.txt 2>&1
if [ "${2:-}" != "skip-tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1
  echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_gpu.log
fi
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
echo "bench ref rc=$?"; cat gpurun_out/${tag}_bench_ref.json
# launch list of the bench command (1 warm-up + 1 timed + 1 instrumented step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/${tag}_launches_256.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline \
    > gpurun_out/${tag}_launches.log 2>&1
echo "launch list rc=$?"
# full capture: the particle kernels and the per-step grid kernels of the second step
PART='k_advect_collide|k_keys_hist|k_scatter_perm|k_sort_within_cells|k_gather|k_p2g_brick|k_build_system|k_apply_pressure|k_correct_tiled|k_extrapolate|k_g2p'
timeout 1500 tools/ncu_capture.sh ${tag}_particles "${PART}" 11 11 --steps 1 --warmup 1
# full capture: PCG + multigrid kernels of one iteration in the second step (first 60 launches after the skip)
PCG='k_spmv|k_update|k_dot|k_xpby|k_mg_|k_pcg'
timeout 900 tools/ncu_capture.sh ${tag}_pcg "${PCG}" 100 60 --steps 1 --warmup 1
ls -la gpurun_out
