#!/bin/bash
# Runs ON the GPU box (under gpurun): parity tests of the current defaults, a memcheck pass over one small golden
# scene, the A/B sweep of kernel variants on the bench scene, and ncu captures of the new kernels.
# Usage: tools/gpu_sweep.sh <tag> [grid]
set -u
tag=${1:-sweep}
grid=${2:-256}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_smi.txt 2>&1
OLD="p2g=1,correct=0,mg_tail=1,warm_start=0,red_blocks=16384"
timeout -k 10 700 python -m pytest tests -m gpu -q --timeout 150 --timeout-method=thread > gpurun_out/${tag}_pytest_gpu.log 2>&1
rc=$?
echo "pytest rc=$rc"; tail -25 gpurun_out/${tag}_pytest_gpu.log
if [ $rc -ne 0 ]; then
  for one in "p2g=0" "correct=1" "mg_tail=0" "warm_start=1"; do
    LFK_TUNE="$OLD,$one" timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -q --timeout 100 --timeout-method=thread \
        -k "golden or variants or fused_step_equals or warm" > gpurun_out/${tag}_pytest_${one%%=*}.log 2>&1
    echo "[$one] rc=$?"; tail -4 gpurun_out/${tag}_pytest_${one%%=*}.log
  done
fi
timeout -k 10 200 compute-sanitizer --tool memcheck --print-limit 30 python -m pytest tests/test_gpu_parity.py -q -x \
    -k "stages_match_golden and dam_break and not 0-" > gpurun_out/${tag}_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/${tag}_memcheck.log | head -12
timeout -k 10 600 python tools/variant_sweep.py --grid ${grid} --tag ${tag} > gpurun_out/${tag}_sweep.log 2>&1
echo "sweep rc=$?"; cat gpurun_out/${tag}_sweep.log | tail -12
NEWK='k_p2g_march|k_correct_tiled2'
timeout -k 10 300 tools/ncu_capture.sh ${tag}_newk "${NEWK}" 2 2 --steps 1 --warmup 1 --grid ${grid}
PCG='k_spmv|k_update|k_dot|k_xpby|k_mg_|k_pcg'
timeout -k 10 300 tools/ncu_capture.sh ${tag}_pcg "${PCG}" 100 70 --steps 1 --warmup 1 --grid ${grid}
ls -la gpurun_out
