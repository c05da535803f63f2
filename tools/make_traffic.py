#!/usr/bin/env python3
"""profiles/traffic.json from `ncu --page raw --csv` dumps: DRAM bytes (read + write) per column and launch
for each kernel.  usage: make_traffic.py WORKLOAD COLUMNS_PER_LAUNCH raw1.csv [raw2.csv ...]"""
import csv, json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
W, ncol = sys.argv[1], int(sys.argv[2])
names = {"lw_prep": "lw_prep", "lw_taumol": "lw_taumol", "lw_rtrn": "lw_rtrn", "sw_prep": "sw_prep", "sw_taumol": "sw_taumol", "sw_solver": "sw_solver", "sw_finish": "sw_solver",
         "lw_column": "lw_column", "lw_finish": "lw_column", "sw_column": "sw_column", "sw_cfinish": "sw_column"}
out_path = os.path.join(ROOT, "profiles", "traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
out[W] = {}        # a capture replaces the workload's entry (kernels that no longer run must not linger)
for f in sys.argv[3:]:
    rows = list(csv.reader(open(f)))
    hdr, units = rows[0], rows[1]
    i = {h: k for k, h in enumerate(hdr)}
    seen = set()
    for r in rows[2:]:
        k = next((v for n, v in names.items() if n in r[i["Kernel Name"]]), None)
        if not k:
            continue
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd = float(r[i["dram__bytes_read.sum"]]) * scale[units[i["dram__bytes_read.sum"]]]
        wr = float(r[i["dram__bytes_write.sum"]]) * scale[units[i["dram__bytes_write.sum"]]]
        dur = float(r[i["gpu__time_duration.sum"]])
        e = {"dram_bytes_per_column": (rd + wr) / ncol, "dram_read_bytes": rd, "dram_write_bytes": wr,
             "columns_per_launch": ncol, "duration_ms_under_ncu": dur,
             "fp64_pipe_pct": float(r[i["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]]),
             "l1tex_throughput_pct": float(r[i["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]]),
             "dram_throughput_pct": float(r[i["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
             "issue_active_pct": float(r[i["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
             "registers_per_thread": int(float(r[i["launch__registers_per_thread"]])), "source": os.path.basename(f),
             "kernels": [r[i["Kernel Name"]].split("(")[0]]}
        if k in seen:          # a step made of two kernels (prep: per cell + per column): add traffic and time,
            p = out[W][k]      # keep the utilisation figures of the longer one
            for key in ("dram_bytes_per_column", "dram_read_bytes", "dram_write_bytes", "duration_ms_under_ncu"):
                e[key] += p[key]
            if p["duration_ms_under_ncu"] > dur:
                for key in ("fp64_pipe_pct", "l1tex_throughput_pct", "dram_throughput_pct", "issue_active_pct", "registers_per_thread"):
                    e[key] = p[key]
            e["kernels"] = p["kernels"] + e["kernels"]
        out[W][k] = e
        seen.add(k)
# stamp: the tree these numbers were profiled on (bench.py reports them only while the CUDA sources are unchanged)
sys.path.insert(0, ROOT)
from mima_b200.build import source_hash  # noqa: E402
import datetime  # noqa: E402
out[W]["_stamp"] = {"csrc_sha256": source_hash(), "when": datetime.datetime.utcnow().strftime("%Y-%m-%dT%H:%MZ"),
                    "how": "ncu --set full --clock-control none, first pass of every kernel (tools/gpu_profile.sh)"}
json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps({k: v for k, v in out[W].items() if k != "_stamp"}, indent=1))
