#!/bin/bash
set -u
mkdir -p gpurun_out
for W in T42L40 T85L40; do
timeout 600 python tools/gpu_sweep.py $W "taumol_bin=0" "taumol_bin=1,taumol_run=1" "taumol_bin=1,taumol_run=2" "taumol_bin=1,taumol_run=8" 2>&1 | tee -a gpurun_out/r2e_sweep.txt
done
for C in 8192 16384 32768; do
timeout 600 python tools/gpu_sweep.py T170L60 "chunk=$C,taumol_bin=0" "chunk=$C,taumol_bin=1,taumol_run=1" "chunk=$C,taumol_bin=1,taumol_run=2" 2>&1 | tee -a gpurun_out/r2e_sweep.txt
done
