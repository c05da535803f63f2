#!/bin/bash
# final tree: GPU suite, ncu captures for the traffic.json stamp, the T170L60 bench line
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
bash tools/gpu_profile.sh T170L60 2>&1 | tail -2
