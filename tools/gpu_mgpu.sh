#!/bin/bash
# Runs ON an N-GPU box (gpurun --gpus N): the multi-GPU check against the single-GPU run, then a short weak-scaling
# bench line WITH the end-to-end leg (host buffers, per-rank particle counts).  Everything lands in gpurun_out/.
# Usage: tools/gpu_mgpu.sh <tag> [nranks] [bench steps]
set -u
tag=${1:-mgpu}
nr=${2:-2}
steps=${3:-5}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
export PYTHONFAULTHANDLER=1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node ${nr} --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
if [ "${MGPU_SKIP_CHECK:-0}" != "1" ]; then
timeout -k 5 ${MGPU_CHECK_TIMEOUT:-200} bash -c "$(declare -f run); nr=${nr}; run 29511 tests/mgpu_check.py" \
    > gpurun_out/${tag}_mgpu_check_${nr}gpu.log 2>&1
echo "mgpu_check rc=$?" | tee -a gpurun_out/${tag}_mgpu_check_${nr}gpu.log
grep -v "^W\|^\*\*\*\|^$" gpurun_out/${tag}_mgpu_check_${nr}gpu.log | tail -12
fi
BENCH_WATCHDOG_S=${MGPU_BENCH_TIMEOUT:-300} timeout -k 5 $((${MGPU_BENCH_TIMEOUT:-300} + 20)) bash -c "$(declare -f run); nr=${nr}; run 29512 bench.py --gpus ${nr} --steps ${steps} --warmup 3 --no-cpu-baseline ${MGPU_BENCH_ARGS:-}" \
    > gpurun_out/${tag}_bench_${nr}gpu.json 2> gpurun_out/${tag}_bench_${nr}gpu.err
echo "bench rc=$?"; cat gpurun_out/${tag}_bench_${nr}gpu.json; grep -v "^W\|^\*\*\*\|^$" gpurun_out/${tag}_bench_${nr}gpu.err | tail -15
