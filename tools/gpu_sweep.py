#!/usr/bin/env python3
"""Developer tool: sweep launch-tuning options and print per-kernel device times (ms per step)."""
import ctypes as C, json, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
W = sys.argv[1] if len(sys.argv) > 1 else "T85L40"
for key, vals in (("sw_solver_pad_kb", [0, 40, 52, 70, 100, 200]), ("lw_rtrn_pad_kb", [0, 20, 35, 52, 90, 200])):
    for v in vals:
        env = dict(os.environ, RRTMG_TUNE=f"{key}={v}")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "2", "--workload", W, "--no-cpu"],
                           capture_output=True, text=True, env=env)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            pk = d["roofline"]["per_kernel"]
            print(key, v, "ms/step=%.2f" % d["ms_per_step"], {k: round(x["ms_per_step"], 2) for k, x in pk.items()}, flush=True)
        except Exception as e:
            print(key, v, "failed", e, r.stderr[-500:], flush=True)
