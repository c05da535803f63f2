#!/bin/bash
# round 2, session f: whole GPU suite on the new tree, bench lines for every workload, reference arm
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for W in T170L60 T85L40 T42L40 T42L40-4xCO2; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $W > gpurun_out/r2f_bench_$W.json 2> gpurun_out/r2f_bench_$W.err || tail -5 gpurun_out/r2f_bench_$W.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2f_bench_$W.json"))
    pk=d["roofline"]["per_kernel"]
    print("$W", "ms/step=%.2f"%d["ms_per_step"], "Mcol/s=%.3f"%(d["value"]/1e6), "step_frac=%.3f"%d["roofline"]["step_frac"], "e2e_ms=%.1f"%d["e2e"]["ms_per_step"], "e2e_all=%.1f"%d["e2e_all_outputs"]["ms_per_step"], "run_rrtmg=%.1f"%d["e2e_run_rrtmg"]["ms_per_step"], {k:round(v["ms_per_step"],2) for k,v in pk.items()}, d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["clocks"])
except Exception as e:
    print("bench $W failed", e); print(open("gpurun_out/r2f_bench_$W.err").read()[-1500:])
PY
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference.json 2>/dev/null; cut -c1-400 gpurun_out/r2f_bench_reference.json
