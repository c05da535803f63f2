#!/bin/bash
# ncu --set full capture of each kernel of one LW+SW step (one launch each), plus the launch list.
set -u
W=${1:-T85L40}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lw_|sw_" -s 12 -c 6 -f -o gpurun_out/prof_${W} \
    python bench.py --steps 1 --warmup 2 --workload $W --no-cpu --chunk 1000000 > gpurun_out/ncu_full_${W}.log 2>&1
tail -3 gpurun_out/ncu_full_${W}.log
ls -la gpurun_out/
