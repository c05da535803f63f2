#!/bin/bash
# round 2, session j: the whole GPU suite + smoke + ncu captures (traffic.json stamp) + bench lines on the final tree
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/gpu_profile.sh T170L60 2>&1 | tail -3
for W in T170L60 T42L40; do
  timeout 600 python bench.py --steps 20 --warmup 3 --workload $W > gpurun_out/r2j_bench_$W.json 2> gpurun_out/r2j_bench_$W.err || tail -5 gpurun_out/r2j_bench_$W.err
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 --workload T42L40 > gpurun_out/r2j_bench_reference_T42L40_full.json 2>/dev/null
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j_bench_reference.json 2>/dev/null
rm -f gpurun_out/*.ncu-rep
python - <<'PY'
import json
for W in ("T170L60", "T42L40"):
    d=json.load(open("gpurun_out/r2j_bench_%s.json"%W))
    print(W, "ms/step=%.2f"%d["ms_per_step"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], d["roofline"]["traffic_source"][:60], "e2e %.1f"%d["e2e"]["ms_per_step"])
for f in ("r2j_bench_reference_T42L40_full.json", "r2j_bench_reference.json"):
    d=json.load(open("gpurun_out/"+f)); print(f, d["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["sample"])
PY
