#!/bin/bash
# round 2, session o: LW taumol without the exactly-zero terms; source-level profile of lw_rtrn
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q -m gpu 2>&1 | tail -2
python tools/gpu_sweep.py T170L60 "" 2>&1 | tee gpurun_out/r2o_sweep.txt
python tools/gpu_sweep.py T42L40-4xCO2 "" 2>&1 | tee -a gpurun_out/r2o_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lw_rtrn -s 2 -c 1 -f -o gpurun_out/r2o_lwrtrn \
     python bench.py --steps 1 --warmup 1 --workload T170L60 --no-cpu > gpurun_out/r2o_lwrtrn.log 2>&1
ncu -i gpurun_out/r2o_lwrtrn.ncu-rep --page source --csv > gpurun_out/r2o_lwrtrn_source.csv 2>/dev/null
rm -f gpurun_out/r2o_lwrtrn.ncu-rep
