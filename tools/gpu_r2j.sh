#!/bin/bash
# final-tree session: the whole GPU suite + smoke + compute-sanitizer + ncu captures (traffic.json stamp) + bench lines
set -u
TAG=${1:-r2fin}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for TOOL in memcheck initcheck; do
  timeout 900 compute-sanitizer --tool $TOOL --print-limit 5 python tools/sanitize_workload.py > gpurun_out/${TAG}_sanitizer_$TOOL.log 2>&1
  echo "compute-sanitizer $TOOL: $(grep -E 'ERROR SUMMARY' gpurun_out/${TAG}_sanitizer_$TOOL.log | tail -1)"
done
bash tools/gpu_profile.sh T170L60 2>&1 | tail -3
for W in T170L60 T85L40 T42L40 T42L40-4xCO2; do
  timeout 600 python bench.py --steps 20 --warmup 3 --workload $W > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err || tail -5 gpurun_out/${TAG}_bench_$W.err
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null
rm -f gpurun_out/*.ncu-rep
python - <<PY
import json
for W in ("T170L60", "T85L40", "T42L40", "T42L40-4xCO2"):
    d=json.load(open("gpurun_out/${TAG}_bench_%s.json"%W))
    print(W, "ms/step=%.2f"%d["ms_per_step"], "step_frac", d["roofline"].get("step_frac"), "dom", d["roofline"].get("kernel"), d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "e2e %.1f"%d["e2e"]["ms_per_step"])
d=json.load(open("gpurun_out/${TAG}_bench_reference.json")); print("reference", d["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["sample"])
PY
