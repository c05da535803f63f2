#!/bin/bash
# round 2, session ab: host pipeline block size for the end-to-end path now that the kernels prefer larger passes
set -u
mkdir -p gpurun_out
: > gpurun_out/r2ab_e2e.txt
for HC in 0 32768 65536; do
  RRTMG_TUNE="host_chunk=$HC" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2ab_hc$HC.json 2>/dev/null
  python - <<PY | tee -a gpurun_out/r2ab_e2e.txt
import json
d=json.load(open("gpurun_out/r2ab_hc$HC.json"))
print("host_chunk=$HC", "device %.2f"%d["ms_per_step"], "e2e %.2f"%d["e2e"]["ms_per_step"], "all_outputs %.2f"%d["e2e_all_outputs"]["ms_per_step"], "run_rrtmg %.2f"%d["e2e_run_rrtmg"]["ms_per_step"])
PY
done
for W in T85L40 T42L40; do for HC in 0 8192 16384; do
  RRTMG_TUNE="host_chunk=$HC" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --workload $W > gpurun_out/r2ab_${W}_hc$HC.json 2>/dev/null
  python - <<PY | tee -a gpurun_out/r2ab_e2e.txt
import json
d=json.load(open("gpurun_out/r2ab_${W}_hc$HC.json"))
print("$W host_chunk=$HC", "device %.2f"%d["ms_per_step"], "e2e %.2f"%d["e2e"]["ms_per_step"], "all_outputs %.2f"%d["e2e_all_outputs"]["ms_per_step"], "run_rrtmg %.2f"%d["e2e_run_rrtmg"]["ms_per_step"])
PY
done; done
