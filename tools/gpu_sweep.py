#!/usr/bin/env python3
"""Developer tool: run bench.py once per tuning setting and print per-kernel device times (ms per step).
usage: gpu_sweep.py WORKLOAD "key=v[,key=v]" ..."""
import json, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
W = sys.argv[1] if len(sys.argv) > 1 else "T85L40"
for setting in (sys.argv[2:] or [""]):
    env = dict(os.environ, RRTMG_TUNE=setting)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--workload", W, "--no-cpu"],
                       capture_output=True, text=True, env=env)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        pk = d["roofline"]["per_kernel"]
        print(W, setting or "(default)", "ms/step=%.2f" % d["ms_per_step"], {k: round(x["ms_per_step"], 2) for k, x in pk.items()}, flush=True)
    except Exception as e:
        print(W, setting, "failed", e, r.stderr[-500:], flush=True)
