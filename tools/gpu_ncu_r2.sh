#!/bin/bash
# Runs ON a 1-GPU box: (1) per-launch time + DRAM traffic of every kernel of one 256^3 bench step (ncu, one pass per
# kernel; also the launch list); (2) a full-set capture with source of the particle kernels at 128^3; (3) a full-set
# capture of the PCG iteration's kernels at 256^3; (4) the projection-only sweep on one GPU.  Outputs in gpurun_out/.
set -u
tag=${1:-r2}
mkdir -p gpurun_out
timeout -k 5 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -c 6000 --csv --log-file gpurun_out/${tag}_traffic_256.csv python bench.py --steps 1 --warmup 1 --no-e2e \
    --no-cpu-baseline > gpurun_out/${tag}_traffic.log 2>&1
echo "traffic rc=$?"; wc -l gpurun_out/${tag}_traffic_256.csv
python tools/ncu_traffic_table.py gpurun_out/${tag}_traffic_256.csv gpurun_out/${tag}_traffic_256.json && head -c 1500 gpurun_out/${tag}_traffic_256.json
timeout -k 5 300 bash tools/ncu_capture.sh ${tag}_particles128 "k_correct_tile|k_p2g_march|k_g2p|k_advect_collide" 8 4 --grid 128 --steps 1 --warmup 3
echo "capture particles rc=$?"
LFK_TUNE=graph=0 timeout -k 5 300 bash tools/ncu_capture.sh ${tag}_pcg256 "k_spmv_dot|k_update_pr|k_mg_rbgs_l0|k_mg_final_l0|k_mg_restrict_l0|k_xpby|k_mg_tail" 60 24 --grid 256 --steps 1 --warmup 2
echo "capture pcg rc=$?"
timeout -k 5 300 python tools/projection_sweep.py --grids 128,256,512,1024 --tag ${tag} > gpurun_out/${tag}_projection.log 2>&1
echo "projection rc=$?"; cut -c1-400 gpurun_out/${tag}_projection.log | tail -6
ls -la gpurun_out | tail -20
