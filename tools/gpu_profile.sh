#!/bin/bash
# ncu captures of one LW+SW step of the bench workload (default chunking, single stream):
#  (1) --set full for the first pass (chunk) of every kernel: 4 SW launches (prep_cell, prep, column, cfinish) + 4 LW launches
#      (prep_cell, prep, column, finish), raw CSV exported;
#  (2) the launch list with device times for two whole steps.
# Outputs under gpurun_out/ (kept below the 64 MiB copy-back limit).
set -u
W=${1:-T170L60}
mkdir -p gpurun_out
export RRTMG_SKIP_NIGHT=1
NPASS=$(python -c "n={'T42L40':8192,'T85L40':32768,'T170L60':131072,'T341L80':524288}['$W']; print((n+131071)//131072)")
STEP=$((NPASS * 8)); SKIP=$((2 * STEP))
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lw_|sw_" -s $SKIP -c 4 -f -o gpurun_out/prof_${W}_sw \
    python bench.py --steps 1 --warmup 2 --workload $W --no-cpu > gpurun_out/ncu_full_${W}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lw_|sw_" -s $((SKIP + NPASS * 4)) -c 4 -f -o gpurun_out/prof_${W}_lw \
    python bench.py --steps 1 --warmup 2 --workload $W --no-cpu >> gpurun_out/ncu_full_${W}.log 2>&1
tail -2 gpurun_out/ncu_full_${W}.log
# the reports themselves (2 x ~40 MB with the source pages of the column kernels) exceed the 64 MiB copy-back limit: export the raw
# pages and the per-instruction source page of the column kernel here, keep only those
for x in sw lw; do
  ncu -i gpurun_out/prof_${W}_$x.ncu-rep --page raw --csv > gpurun_out/prof_${W}_${x}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_${W}_$x.ncu-rep --page source --csv -k regex:"${x}_column" > gpurun_out/prof_${W}_${x}col_source.csv 2>/dev/null
  gzip -f gpurun_out/prof_${W}_${x}col_source.csv
  rm -f gpurun_out/prof_${W}_$x.ncu-rep
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lw_|sw_" -s $SKIP -c $((2 * STEP)) --csv --log-file gpurun_out/launches_${W}.csv \
    python bench.py --steps 2 --warmup 2 --workload $W --no-cpu > gpurun_out/ncu_list_${W}.log 2>&1
ls -la gpurun_out/ | tail -8
