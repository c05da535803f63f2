#!/bin/bash
# adaptive host blocks: e2e of the small workloads on one GPU, new default against one block of 16384
set -u
for W in T42L40 T85L40 T170L60; do
  for T in "" "host_chunk=16384,run_chunk=16384"; do
    RRTMG_TUNE="$T" timeout 600 python bench.py --steps 10 --warmup 3 --workload $W --no-cpu > gpurun_out/hc_$W.json 2>/dev/null
    python - <<PY
import json
d=json.load(open("gpurun_out/hc_$W.json"))
print("$W", "[$T]", "dev %.2f"%d["ms_per_step"], "e2e %.2f"%d["e2e"]["ms_per_step"], "e2e_all %.2f"%d["e2e_all_outputs"]["ms_per_step"], "run_rrtmg %.2f"%d["e2e_run_rrtmg"]["ms_per_step"])
PY
  done
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
