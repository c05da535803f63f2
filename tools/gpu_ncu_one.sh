#!/bin/bash
# ncu --set full of one launch of kernels matching $1 under tuning $2; tag $3.  Outputs raw + source CSV under gpurun_out/.
set -u
K=$1; T=$2; TAG=$3
mkdir -p gpurun_out
RRTMG_TUNE="$T" timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 4 -c 1 -f -o gpurun_out/ncu_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
ncu -i gpurun_out/ncu_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_$TAG.ncu-rep --page source --csv > gpurun_out/ncu_${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/ncu_$TAG*
