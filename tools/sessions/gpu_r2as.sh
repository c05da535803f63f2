#!/bin/bash
# round 2, session as: the default bench line of the final tree (4 pipeline slots), then 5 and 6 slots
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r2fin5_bench_T170L60.json 2> gpurun_out/r2fin5_bench.err
O=gpurun_out/r2as_sweep.txt
: > $O
for S in 5 6; do
  RRTMG_TUNE=host_slots=$S python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('host_slots=$S', 'device %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], 'e2e_all %.2f'%d['e2e_all_outputs']['ms_per_step'], 'run_rrtmg %.2f'%d['e2e_run_rrtmg']['ms_per_step'])" | tee -a $O
done
python -c "
import json
d=json.load(open('gpurun_out/r2fin5_bench_T170L60.json'))
print('default', 'device %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], 'e2e_all %.2f'%d['e2e_all_outputs']['ms_per_step'], 'run_rrtmg %.2f'%d['e2e_run_rrtmg']['ms_per_step'], d['roofline']['traffic'])" | tee -a $O
