#!/bin/bash
# GPU session: parity for every solver variant, then bench lines per variant (T170L60).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run_bench() {   # tag, tune
  RRTMG_TUNE="$2" timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$1.json"))
    pk=d["roofline"]["per_kernel"]
    print("$1", "ms/step=%.2f"%d["ms_per_step"], "e2e_ms=%.1f"%d["e2e"]["ms_per_step"], {k:round(v["ms_per_step"],2) for k,v in pk.items()}, d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("bench $1 failed", e); print(open("gpurun_out/bench_$1.err").read()[-1500:])
PY
}
for T in "lw_rtrn_variant=1,sw_solver_variant=0" "lw_rtrn_variant=0,sw_solver_variant=1" "lw_rtrn_variant=1,sw_solver_variant=2" "lw_rtrn_variant=1,sw_solver_variant=3"; do
  echo "== pytest $T"
  RRTMG_TUNE="$T" timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
done
run_bench base "lw_rtrn_variant=0,sw_solver_variant=0"
run_bench lw1 "lw_rtrn_variant=1,sw_solver_variant=0"
run_bench lw1sw1 "lw_rtrn_variant=1,sw_solver_variant=1"
run_bench lw1sw2 "lw_rtrn_variant=1,sw_solver_variant=2"
run_bench lw1sw3 "lw_rtrn_variant=1,sw_solver_variant=3"
