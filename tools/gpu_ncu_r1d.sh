#!/bin/bash
# Runs ON the GPU box: (1) per-launch time + DRAM traffic of every kernel of one 256^3 bench step (one ncu pass per
# kernel: the `traffic` figures of bench.py's roofline object); (2) a full-set capture with source of the three
# particle kernels at 128^3 (same per-particle behaviour, 8x less replay time).  Outputs in gpurun_out/.
set -u
tag=${1:-r1d}
mkdir -p gpurun_out
timeout -k 5 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -c 4000 --csv --log-file gpurun_out/${tag}_traffic_256.csv python bench.py --steps 1 --warmup 1 --no-e2e \
    --no-cpu-baseline > gpurun_out/${tag}_traffic.log 2>&1
echo "traffic rc=$?"; wc -l gpurun_out/${tag}_traffic_256.csv
timeout -k 5 240 bash tools/ncu_capture.sh ${tag}_particles128 "k_correct_tiled3|k_p2g_march|k_g2p" 6 3 --grid 128 --steps 1 --warmup 3
echo "capture rc=$?"
