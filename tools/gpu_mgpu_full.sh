#!/bin/bash
# Runs ON an N-GPU box: multi-GPU check, the bench line WITH the end-to-end leg (what the driver runs), and the
# projection-only sweep (BASELINE configs[4]) at N ranks.  Usage: tools/gpu_mgpu_full.sh <tag> <nranks> [grids]
set -u
tag=$1; nr=$2; grids=${3:-128,256,512,1024}
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1 NCCL_DEBUG=WARN
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node ${nr} --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout -k 5 300 bash -c "$(declare -f run); nr=${nr}; run 29511 tests/mgpu_check.py" > gpurun_out/${tag}_mgpu_check_${nr}gpu.log 2>&1
echo "mgpu_check rc=$?" | tee -a gpurun_out/${tag}_mgpu_check_${nr}gpu.log
grep "FAIL\|mgpu_check\|method" gpurun_out/${tag}_mgpu_check_${nr}gpu.log | cut -c1-300 | head -12
BENCH_WATCHDOG_S=400 timeout -k 5 420 bash -c "$(declare -f run); nr=${nr}; run 29512 bench.py --gpus ${nr} --steps 10 --warmup 3 --no-cpu-baseline" \
    > gpurun_out/${tag}_bench_${nr}gpu.json 2> gpurun_out/${tag}_bench_${nr}gpu.err
echo "bench rc=$?"; cut -c1-1200 gpurun_out/${tag}_bench_${nr}gpu.json; grep -v "^W\|^\*\*\*\|^$\|OMP_NUM" gpurun_out/${tag}_bench_${nr}gpu.err | tail -5
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${nr}gpu.json").read().strip().splitlines()[-1])
    print("  ms/step %.2f value %.4g iters/step %.1f pcg_it_ms %.3f e2e %s" % (d["ms_per_step"], d["value"], d["config"]["pcg_iters_per_step"], d["roofline"]["all"]["pcg_iteration"]["ms"], {k: (v if not isinstance(v, dict) else v.get("value")) for k, v in (d["e2e"] or {}).items() if k in ("value", "positions_streaming")}))
    print("  phases", {k: round(v,2) for k,v in d["phase_ms"].items()})
except Exception as ex:
    print("  no line:", ex)
PY
timeout -k 5 400 bash -c "$(declare -f run); nr=${nr}; run 29513 tools/projection_sweep.py --grids ${grids} --tag ${tag}" > gpurun_out/${tag}_projection_${nr}gpu.log 2>&1
echo "projection rc=$?"; grep '^{' gpurun_out/${tag}_projection_${nr}gpu.log | cut -c1-330
