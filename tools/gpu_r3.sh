#!/bin/bash
set -u
mkdir -p gpurun_out
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
for T in $PYTEST_TUNES; do
  echo "== pytest $T"
  RRTMG_TUNE="$T" timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
done
i=0
for T in $BENCH_TUNES; do
  run_bench v$i "$T"; i=$((i+1))
done
