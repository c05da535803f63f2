#!/bin/bash
# round 2, session ar: slots of the host-pointer pipelines (2 as before, 3, 4) on the end-to-end numbers; then the captures for the
# traffic.json stamp and the bench line of the final tree
set -u
mkdir -p gpurun_out
O=gpurun_out/r2ar_sweep.txt
: > $O
for S in 2 3 4; do
  RRTMG_TUNE=host_slots=$S python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('host_slots=$S', 'device %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], 'e2e_all %.2f'%d['e2e_all_outputs']['ms_per_step'], 'run_rrtmg %.2f'%d['e2e_run_rrtmg']['ms_per_step'])" | tee -a $O
done
for S in 2 3; do
  RRTMG_TUNE=host_slots=$S python bench.py --steps 5 --warmup 3 --no-cpu --workload T42L40 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('T42L40 host_slots=$S', 'device %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], 'e2e_all %.2f'%d['e2e_all_outputs']['ms_per_step'], 'run_rrtmg %.2f'%d['e2e_run_rrtmg']['ms_per_step'])" | tee -a $O
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_run_rrtmg.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2 | tee -a $O
bash tools/gpu_profile.sh T170L60 2>&1 | tail -2
