#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, ncu launch list.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_T170L60.json 2> gpurun_out/bench_T170L60.err; echo "bench rc=$?"
cat gpurun_out/bench_T170L60.json; tail -5 gpurun_out/bench_T170L60.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload T85L40 --no-cpu > gpurun_out/bench_T85L40.json 2> gpurun_out/bench_T85L40.err
cat gpurun_out/bench_T85L40.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
