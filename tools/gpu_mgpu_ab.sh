#!/bin/bash
# Runs ON an N-GPU box: multi-GPU check, then short bench lines (device-timed only) for A/B settings of LFK_TUNE.
# Usage: tools/gpu_mgpu_ab.sh <tag> <nranks> "<tune1>" "<tune2>" ...   ("-" = library defaults)
set -u
tag=$1; nr=$2; shift 2
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1 NCCL_DEBUG=WARN
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node ${nr} --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
if [ "${MGPU_SKIP_CHECK:-0}" != "1" ]; then
timeout -k 5 240 bash -c "$(declare -f run); nr=${nr}; run 29511 tests/mgpu_check.py" > gpurun_out/${tag}_mgpu_check_${nr}gpu.log 2>&1
echo "mgpu_check rc=$?" | tee -a gpurun_out/${tag}_mgpu_check_${nr}gpu.log
grep "FAIL\|mgpu_check\|method" gpurun_out/${tag}_mgpu_check_${nr}gpu.log | cut -c1-300 | head -12
fi
k=0
for tune in "$@"; do
  k=$((k+1)); name=$(echo "$tune" | tr '=,' '__'); [ "$tune" = "-" ] && tune=""
  LFK_TUNE="$tune" BENCH_WATCHDOG_S=200 timeout -k 5 230 bash -c "$(declare -f run); nr=${nr}; run $((29520+k)) bench.py --gpus ${nr} --steps ${MGPU_STEPS:-5} --warmup 3 --no-cpu-baseline --no-e2e" \
    > gpurun_out/${tag}_bench_${nr}gpu_${name}.json 2> gpurun_out/${tag}_bench_${nr}gpu_${name}.err
  echo "bench[$tune] rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${nr}gpu_${name}.json").read().strip().splitlines()[-1])
    print("  ms/step %.2f value %.4g iters/step %.1f phases %s pcg_it_ms %.3f" % (d["ms_per_step"], d["value"], d["config"]["pcg_iters_per_step"], {k: round(v,2) for k,v in d["phase_ms"].items()}, d["roofline"]["all"]["pcg_iteration"]["ms"]))
except Exception as ex:
    print("  no line:", ex)
PY
  grep -v "^W\|^\*\*\*\|^$\|OMP_NUM" gpurun_out/${tag}_bench_${nr}gpu_${name}.err | tail -4
done
