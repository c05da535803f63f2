#!/bin/bash
# round 2, session h (8 GPUs): strong scaling of the one T170L60 batch (the headline the 85 % target is quoted on), C5
set -u
mkdir -p gpurun_out
N=${1:-8}
run() { # workload extra
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload $1 $2 > gpurun_out/r2h_bench_${1}_${N}gpu.json 2> gpurun_out/r2h_bench_${1}_${N}gpu.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2h_bench_${1}_${N}gpu.json").read().strip().splitlines()[-1])
    print("$1 N=$N", d["scaling"], "ms/step=%.2f"%d["ms_per_step"], "Mcol/s=%.3f"%(d["value"]/1e6), "step_frac=%.3f"%d["roofline"]["step_frac"], "e2e_ms=%.1f"%d["e2e"]["ms_per_step"], "e2e_all=%.1f"%d["e2e_all_outputs"]["ms_per_step"], "run_rrtmg=%.1f"%d["e2e_run_rrtmg"]["ms_per_step"], "weak:", d.get("weak_scaling"))
except Exception as e:
    print("bench $1 failed", e); print(open("gpurun_out/r2h_bench_${1}_${N}gpu.err").read()[-1500:])
PY
}
run T170L60 "--no-cpu"
run T341L80 "--no-cpu"
