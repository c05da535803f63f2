#!/bin/bash
# round 2, session a: parity of the binned taumol + L2-stack SW solver, then timing sweeps
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8
W=T170L60
timeout 1500 python tools/gpu_sweep.py $W \
  "taumol_bin=0" \
  "taumol_bin=1,taumol_order=0" \
  "taumol_bin=1,taumol_order=1" \
  "sw_solver_variant=5,x0=28,x1=0,x2=1" \
  "sw_solver_variant=5,x0=28,x1=3,x2=1" \
  "sw_solver_variant=5,x0=28,x1=3" \
  "sw_solver_variant=5,x0=24,x1=3" \
  "sw_solver_variant=5,x0=24,x1=1" \
  "sw_solver_variant=5,x0=24,x1=2" \
  "sw_solver_variant=5,x0=20,x1=3" \
  "sw_solver_variant=5,x0=20,x1=3,x2=1" \
  "sw_solver_variant=5,x0=20,x1=0" \
  "sw_solver_variant=5,x0=16,x1=3" \
  "sw_solver_variant=5,x0=16,x1=0" \
  "sw_solver_variant=5,x0=12,x1=3" \
  2>&1 | tee gpurun_out/r2a_sweep.txt
