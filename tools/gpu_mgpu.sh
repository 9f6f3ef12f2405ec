#!/bin/bash
# Runs ON a 2-GPU box (gpurun --gpus 2): the multi-GPU check against the single-GPU run, then a short weak-scaling
# bench line.  Everything lands in gpurun_out/.  Usage: tools/gpu_mgpu.sh <tag> [nranks]
set -u
tag=${1:-mgpu}
nr=${2:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
export PYTHONFAULTHANDLER=1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node ${nr} --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout -k 5 ${MGPU_CHECK_TIMEOUT:-110} bash -c "$(declare -f run); nr=${nr}; run 29511 tests/mgpu_check.py" \
    > gpurun_out/${tag}_mgpu_check.log 2>&1
echo "mgpu_check rc=$?"; grep -v "^W\|^\*\*\*\|^$" gpurun_out/${tag}_mgpu_check.log | tail -12
timeout -k 5 ${MGPU_BENCH_TIMEOUT:-90} bash -c "$(declare -f run); nr=${nr}; run 29512 bench.py --gpus ${nr} --steps 3 --warmup 3 --no-e2e --no-cpu-baseline" \
    > gpurun_out/${tag}_bench_${nr}gpu.json 2> gpurun_out/${tag}_bench_${nr}gpu.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench_${nr}gpu.json; tail -5 gpurun_out/${tag}_bench_${nr}gpu.err
