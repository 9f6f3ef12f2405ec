#!/bin/bash
# Runs ON the GPU box (under gpurun): ncu captures of bench.py, exported to small CSVs so that gpurun_out/ stays
# below the 64 MiB copy-back limit.  Usage: tools/ncu_capture.sh <tag> <kernel-regex> <skip> <count> [bench args...]
set -u
tag=$1; regex=$2; skip=$3; count=$4; shift 4
mkdir -p gpurun_out
rep=/tmp/${tag}.ncu-rep
ncu --set full --clock-control none --import-source on -k "regex:${regex}" -s "${skip}" -c "${count}" -f -o /tmp/${tag} \
    python bench.py --no-e2e --no-cpu-baseline "$@" > gpurun_out/${tag}.log 2>&1
ncu -i ${rep} --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i ${rep} --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
ls -la ${rep} gpurun_out/${tag}_raw.csv gpurun_out/${tag}_source.csv
sz=$(stat -c %s ${rep})
if [ "$sz" -lt 20000000 ]; then cp ${rep} gpurun_out/; fi
