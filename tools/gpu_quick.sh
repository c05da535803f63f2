#!/bin/bash
# quick GPU check: parity tests + one bench line per workload given
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
for W in "$@"; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $W --no-cpu > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$W.json"))
    pk=d["roofline"]["per_kernel"]
    print("$W", "ms/step=%.2f"%d["ms_per_step"], "Mcol/s=%.3f"%(d["value"]/1e6), "step_frac=%.3f"%d["roofline"]["step_frac"], "e2e_ms=%.1f"%d["e2e"]["ms_per_step"], {k:round(v["ms_per_step"],2) for k,v in pk.items()}, d["clocks"])
except Exception as e:
    print("bench $W failed", e); print(open("gpurun_out/bench_$W.err").read()[-2000:])
PY
done
