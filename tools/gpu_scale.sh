#!/bin/bash
# strong-scaling line of bench.py at N GPUs (as the driver launches it) + the reference arm; usage: gpu_scale.sh N [workload]
set -u
mkdir -p gpurun_out
N=${1:-8}; W=${2:-T170L60}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 --workload $W --no-cpu > gpurun_out/scale_${W}_${N}gpu.json 2> gpurun_out/scale_${W}_${N}gpu.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --workload $W --no-cpu > gpurun_out/scale_${W}_${N}gpu.json 2> gpurun_out/scale_${W}_${N}gpu.err
fi
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${W}_${N}gpu.json").read().strip().splitlines()[-1])
    ws=d.get("weak_scaling") or {}
    print("$W N=$N", d["scaling"], "ms/step=%.3f"%d["ms_per_step"], "Mcol/s=%.3f"%(d["value"]/1e6), "e2e_ms=%.2f"%d["e2e"]["ms_per_step"], "e2e_all=%.2f"%d["e2e_all_outputs"]["ms_per_step"], "run_rrtmg=%.2f"%d["e2e_run_rrtmg"]["ms_per_step"],
          "| weak: ms=%s Mcol/s=%s e2e_ms=%s"%(ws.get("ms_per_step"), (ws.get("value") or 0)/1e6, (ws.get("e2e") or {}).get("ms_per_step")), d["clocks"]["reasons"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/scale_${W}_${N}gpu.err").read()[-1500:])
PY
