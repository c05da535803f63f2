#!/usr/bin/env python3
"""bench.py -- RRTMG LW+SW columns/second on B200 (metric of BASELINE.json), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload T170L60] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path -- rrtmg_sw + rrtmg_lw, as run_rrtmg calls them
(rrtm_radiation.f90:686-748) -- over one batch of seeded synthetic columns (SURVEY.md 8d generator).

  value      whole-job columns/s with the inputs resident in HBM (device-pointer ABI, CUDA events on the
             launching stream, max over ranks).
  e2e        the same metric through the host-pointer C ABI, called the way the Fortran shim calls it: rrtmg_b200_sw then
             rrtmg_b200_lw on the same pinned host arrays, option share_inputs = 1 (the eleven arrays both codes read cross
             PCIe once), NULL for the clear-sky outputs run_rrtmg never reads; H2D + kernels + D2H every step.
             `e2e_all_outputs` is the same without either economy (every array uploaded twice, all twelve outputs back).
  roofline   dominant kernel: its share of the algorithmic bytes B_alg(L) = 8*(1060 L + 35) B/column
             (per-kernel split in DESIGN.md) divided by its CUDA-event duration, against the
             measured HBM copy bandwidth in MEASURED_PEAKS.json.  `step_frac` is the same for the whole step.
             `traffic` (DRAM bytes per launch) and `fp64_pipe_pct` come from the committed ncu capture
             (profiles/traffic.json) and are reported only while its source stamp matches the CUDA sources in the tree.
  cpu_baseline  the reference's own CPU code -- oracle/_ref, its RRTMG sources machine-translated F90 -> C and compiled
             with gcc (kind "reference"; the image has no Fortran compiler), else the hand-written C port (kind "port")
             -- on all host cores, on a bounded column sample of the same workload.

Multi-GPU: columns are independent, each rank owns a block of latitude rows and drives one GPU; no
collective on the data path (the only communication is the timing reduction).  STRONG scaling is the headline
(BASELINE.json: "T170L60 sharded by latitude rows at 1/2/4/8 B200"): the ONE batch of the named resolution is cut into
N blocks of latitude rows, `value` = its columns / the slowest rank's time.  `weak_scaling` (extra key; --weak makes
it the headline) gives every rank a full batch.

--impl reference times that CPU implementation (all host threads) on the same workload, metric and config; rank 0 only.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# lw_column: the fused clear-sky LW kernel (taumol + rtrn per column) + lw_finish; lw_taumol / lw_rtrn only run in the staged path
KERNELS = ["lw_prep", "lw_taumol", "lw_rtrn", "sw_prep", "sw_taumol", "sw_solver", "lw_column", "sw_column"]
# BASELINE.json configs: C1 T42L40, C2 T85L40, C3 T170L60 (default), C4 T42L40-4xCO2, C5 T341L80
WORKLOADS = {"T42L40": ("T42L40", {}), "T85L40": ("T85L40", {}), "T170L60": ("T170L60", {}), "T341L80": ("T341L80", {}),
             "T42L40-4xCO2": ("T42L40", dict(co2_ppmv=1560.0, ozone="file", secondary_gases=True))}
CPU_KINDS = {"reference": "oracle/_ref: the reference's own RRTMG sources machine-translated F90 -> C (tools/f90_to_c.py) and compiled with "
                          "gcc -O2 -ffp-contract=off; the image has no Fortran compiler",
             "port": "oracle/: the hand-written C restatement (bit-identical to oracle/_ref; used when the translated library is absent)"}


def workload_spec(name: str):
    if name not in WORKLOADS:
        raise SystemExit("bench.py: unknown workload %s (one of %s)" % (name, ", ".join(WORKLOADS)))
    return WORKLOADS[name]


def kernel_alg_bytes(L: int) -> dict:
    """Split of B_alg(L) = 8*(1060 L + 35) bytes/column over the six kernels."""
    return {
        "lw_prep": 8 * (30 * L + 19),          # LW interface inputs actually dereferenced
        "lw_taumol": 8 * (280 * L),            # write taug, fracs (140 g)
        "lw_rtrn": 8 * (280 * L + 6 * L + 4),  # read them back + LW outputs
        "sw_prep": 8 * (10 * L + 8),
        "sw_taumol": 8 * (224 * L),            # write taug, taur (112 g)
        "sw_solver": 8 * (224 * L + 6 * L + 4),
        "lw_column": 8 * (280 * L) + 8 * (280 * L + 6 * L + 4),   # the staging charge of lw_taumol + lw_rtrn, and the outputs
        "sw_column": 8 * (224 * L) + 8 * (224 * L + 6 * L + 4),   # the same for sw_taumol + sw_solver
    }


def b_alg(L: int) -> int:
    return 8 * (1060 * L + 35)


def profiled(workload: str):
    """The committed ncu capture of this workload (profiles/traffic.json) if it was taken on the CUDA sources in the tree
    (stamp = sha256 of the kernel translation units and device headers under mima_b200/csrc, mima_b200.build.source_hash,
    written by tools/make_traffic.py), else None and why."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))[workload]
    except Exception:
        return None, "no ncu capture of this workload under profiles/"
    from mima_b200.build import source_hash
    st = t.get("_stamp", {})
    if st.get("csrc_sha256") != source_hash():
        return None, "profiles/traffic.json was captured on other kernel sources (stamp %s)" % str(st.get("csrc_sha256"))[:12]
    return t, "profiles/traffic.json, ncu --set full of %s (dram__bytes_read.sum + dram__bytes_write.sum; sm__pipe_fp64_cycles_active)" % st.get("when")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled during the timed region (NVML; nvidia-smi as a fallback)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except Exception:
                pass
        return index

    def _sample_nvml(self):
        n = self.nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for nm, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                self.reasons.add(nm)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.samples.append(float(out[0]))
        self.max_mhz = float(out[1])
        for nm, v in zip(names, out[2:]):
            if v.strip().lower().startswith("active"):
                self.reasons.add(nm)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nvml else 0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "source": "nvml" if self.nvml else "nvidia-smi"}


def cpu_arm():
    """(object with rrtmg_sw / rrtmg_lw, kind): the translated reference when its library is there, else the C port."""
    try:
        from oracle import pyref
        if pyref.available():
            return pyref.Reference(), "reference"
    except Exception as e:                          # noqa: BLE001 -- fall through to the port, say why
        sys.stderr.write("bench.py: oracle/_ref unavailable (%s); CPU arm = oracle port\n" % e)
    from oracle.pyoracle import Oracle
    return Oracle(), "port"


def cpu_reference_rate(workload: str, sample_cols: int, repeats: int, threads: int | None = None):
    """columns/s of the reference's CPU implementation on a bounded sample of the workload (mid-latitude rows)."""
    from mima_b200.columns import RESOLUTIONS, make_columns
    base, colkw = workload_spec(workload)
    nlon, nlat, nlay = RESOLUTIONS[base]
    rows = max(1, min(nlat, sample_cols // nlon))
    j0 = (nlat - rows) // 2
    cols = make_columns(base, lat_rows=(j0, j0 + rows), **colkw)
    arm, kind = cpu_arm()
    # all host cores this process may run on (torchrun exports OMP_NUM_THREADS=1, which must not shrink the CPU arm)
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    nt = threads or max(arm.max_threads, ncpu)
    best = None
    times = []
    for _ in range(repeats):
        t = time.perf_counter()
        arm.rrtmg_sw(cols, nthreads=nt)
        arm.rrtmg_lw(cols, nthreads=nt)
        dt = time.perf_counter() - t
        times.append(dt)
        best = dt if best is None else min(best, dt)
    return cols.ncol / best, nt, cols.ncol, times, kind


def config_dict(workload: str, world: int, strong: bool, streams: int = 1):
    """`config` of the JSON line: the same dict from both arms."""
    from mima_b200.columns import RESOLUTIONS
    base, colkw = workload_spec(workload)
    nlon, nlat, nlay = RESOLUTIONS[base]
    total = nlon * nlat * (1 if strong else world)
    return {"workload": workload, "columns_total": total, "columns_per_gpu": total // world, "layers": nlay, "grid": f"{nlon}x{nlat}",
            "sharding": "latitude-row blocks of one batch, one rank per GPU, no collective" if strong
                        else "one full batch per rank, no collective",
            "l2": "inputs + staging per step exceed the 126 MB L2 (no flush needed)",
            "streams": streams, "all_sunlit": True,
            "lw_tables": "synthetic (reference LW k_g file stripped)", "sw_tables": "reference", **colkw}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on all host threads; each step = one pass over a
    bounded sample of the workload (the whole batch would take minutes per step).  Rank 0 only."""
    if rank != 0:
        return
    times_all = []
    sample = nt = kind = None
    for i in range(args.warmup + args.steps):
        _, nt, sample, times, kind = cpu_reference_rate(args.workload, args.cpu_sample, 1)
        if i >= args.warmup:
            times_all.append(times[0])
    total = sum(times_all)
    rate = sample * len(times_all) / total
    cfg = config_dict(args.workload, world, not args.weak)
    line = {
        "impl": "reference", "metric": "RRTMG LW+SW columns/sec", "value": rate, "unit": "columns/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times_all),
        "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": rate, "unit": "columns/s", "cores": nt, "kind": kind, "what": CPU_KINDS[kind],
                         "sample": f"{sample} columns x {cfg['layers']} layers of {args.workload} per step (mid-latitude rows), threads over column blocks"},
        "e2e": {"value": rate, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="T170L60")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=8192, help="columns in the CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--weak", action="store_true",
                    help="weak scaling as the headline: one full batch of the named resolution per rank (default: strong "
                         "scaling, the one batch split by latitude rows over the ranks, e.g. T170L60 on 8 GPUs = 16384 columns per GPU)")
    ap.add_argument("--split", action="store_true", help="(accepted for compatibility: strong scaling is the default)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from mima_b200 import rrtmg, sharding
    from mima_b200.columns import RESOLUTIONS, make_columns

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner ("NCCL version ...", printed to stdout when the
        # environment sets NCCL_DEBUG=VERSION) goes to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    rrtmg.set_device(local_rank)
    rrtmg.rrtmg_lw_ini(allow_synthetic_lw=True)
    rrtmg.rrtmg_sw_ini()
    if args.chunk:
        rrtmg.set_option("chunk", args.chunk)
    # (developer knob RRTMG_TUNE="key=value[,...]" is applied by rrtmg.lib() at load time)
    L_ = rrtmg.lib()

    base, colkw = workload_spec(args.workload)
    nlon, nlat_full, nlay = RESOLUTIONS[base]
    strong = not args.weak
    if strong and nlat_full % world:
        raise SystemExit(f"bench.py: {nlat_full} latitude rows do not divide over {world} ranks")

    def batch(rows=None, seed=20240917):
        return make_columns(base, seed=seed, lat_rows=rows, **colkw)

    # strong scaling (headline): rank r owns latitude rows [r, r+1) * nlat/world of the ONE batch -- a contiguous column
    # range (rrtm_radiation.f90:652: longitude fastest), exactly MiMA's decomposition (spec_mpp.f90:42-49)
    if strong:
        rows = nlat_full // world
        cols = batch((rank * rows, (rank + 1) * rows))
    else:
        cols = batch(seed=20240917 + rank)
    ncol = cols.ncol
    total_cols = nlon * nlat_full if strong else ncol * world
    secondary = bool(colkw.get("secondary_gases"))

    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).pin_memory()

    names = ["play", "tlay", "h2o", "o3", "co2", "plev", "tlev", "tsfc", "albedo", "coszen"]
    if secondary:
        names += ["ch4", "n2o", "o2", "cfc11", "cfc12", "cfc22", "ccl4"]

    class Dev:
        """Device-resident inputs and outputs of one column set."""
        def __init__(self, c):
            self.c, self.n = c, c.ncol
            self.host = {k: pin(getattr(c, k)) for k in names}
            self.d = {k: v.to(dev, non_blocking=True) for k, v in self.host.items()}
            nl, nv = self.n * nlay, self.n * (nlay + 1)
            self.out = {k: torch.empty(nv if "flx" in k else nl, dtype=torch.float64, device=dev)
                        for k in ("lw_uflx", "lw_dflx", "lw_hr", "lw_uflxc", "lw_dflxc", "lw_hrc",
                                  "sw_uflx", "sw_dflx", "sw_hr", "sw_uflxc", "sw_dflxc", "sw_hrc")}

    D = Dev(cols)
    nl, nv = ncol * nlay, ncol * (nlay + 1)
    houts_main = {k: torch.empty(v.numel(), dtype=torch.float64).pin_memory() for k, v in D.out.items()}
    torch.cuda.synchronize()

    def P(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    icld, iaer = C.c_int(0), C.c_int(0)
    stream = torch.cuda.current_stream()
    sh = C.c_void_p(stream.cuda_stream)
    NULL = None
    # RRTMG_TWO_STREAMS=1: SW on a side stream, forked from / joined to the timing stream each step (the two calls are
    # independent, SURVEY.md section 8b "threading").  The timed region defaults to ONE stream so that the per-kernel
    # CUDA-event times behind `roofline` are not inflated by co-running kernels; the two-stream time is reported as
    # `two_streams_ms_per_step` from a short extra pass.
    two_streams = os.environ.get("RRTMG_TWO_STREAMS", "0") == "1"
    side = torch.cuda.Stream(device=dev) if two_streams else None
    sh_sw = C.c_void_p(side.cuda_stream) if two_streams else sh

    def step_device(X=None):
        X = D if X is None else X
        d, o, c = X.d, X.out, X.c
        g = (lambda k: P(d[k])) if secondary else (lambda k: NULL)
        if two_streams:
            side.wait_stream(stream)
        rc = L_.rrtmg_b200_sw_device(C.c_int(X.n), C.c_int(nlay), C.byref(icld), C.byref(iaer),
                                     P(d["play"]), P(d["plev"]), P(d["tlay"]), P(d["tlev"]), P(d["tsfc"]),
                                     P(d["h2o"]), P(d["o3"]), P(d["co2"]), g("ch4"), g("n2o"), g("o2"),
                                     P(d["albedo"]), P(d["albedo"]), P(d["albedo"]), P(d["albedo"]), P(d["coszen"]),
                                     C.c_double(c.adjes), C.c_int(c.dyofyr), C.c_double(c.scon),
                                     C.c_int(0), C.c_int(0), C.c_int(0), *([NULL] * 13),
                                     P(o["sw_uflx"]), P(o["sw_dflx"]), P(o["sw_hr"]), P(o["sw_uflxc"]),
                                     P(o["sw_dflxc"]), P(o["sw_hrc"]), sh_sw)
        if rc:
            raise RuntimeError(L_.rrtmg_b200_last_error().decode())
        rc = L_.rrtmg_b200_lw_device(C.c_int(X.n), C.c_int(nlay), C.byref(icld), C.c_int(0),
                                     P(d["play"]), P(d["plev"]), P(d["tlay"]), P(d["tlev"]), P(d["tsfc"]),
                                     P(d["h2o"]), P(d["o3"]), P(d["co2"]), g("ch4"), g("n2o"), g("o2"),
                                     g("cfc11"), g("cfc12"), g("cfc22"), g("ccl4"), NULL,
                                     C.c_int(0), C.c_int(0), C.c_int(0), *([NULL] * 6), NULL,
                                     P(o["lw_uflx"]), P(o["lw_dflx"]), P(o["lw_hr"]), P(o["lw_uflxc"]),
                                     P(o["lw_dflxc"]), P(o["lw_hrc"]), NULL, NULL, sh)
        if rc:
            raise RuntimeError(L_.rrtmg_b200_last_error().decode())
        if two_streams:
            stream.wait_stream(side)

    def step_host(clear_sky, X=None, houts_x=None):
        X = D if X is None else X
        h = X.host
        ncol = X.n
        houts = houts_main if houts_x is None else houts_x
        g = (lambda k: P(h[k])) if secondary else (lambda k: NULL)
        cs = (lambda t: P(t)) if clear_sky else (lambda t: NULL)
        rc = L_.rrtmg_b200_sw(C.c_int(ncol), C.c_int(nlay), C.byref(icld), C.byref(iaer),
                              P(h["play"]), P(h["plev"]), P(h["tlay"]), P(h["tlev"]), P(h["tsfc"]),
                              P(h["h2o"]), P(h["o3"]), P(h["co2"]), g("ch4"), g("n2o"), g("o2"),
                              P(h["albedo"]), P(h["albedo"]), P(h["albedo"]), P(h["albedo"]), P(h["coszen"]),
                              C.c_double(cols.adjes), C.c_int(cols.dyofyr), C.c_double(cols.scon),
                              C.c_int(0), C.c_int(0), C.c_int(0), *([NULL] * 13),
                              P(houts["sw_uflx"]), P(houts["sw_dflx"]), P(houts["sw_hr"]), cs(houts["sw_uflxc"]),
                              cs(houts["sw_dflxc"]), cs(houts["sw_hrc"]))
        if rc:
            raise RuntimeError(L_.rrtmg_b200_last_error().decode())
        rc = L_.rrtmg_b200_lw(C.c_int(ncol), C.c_int(nlay), C.byref(icld), C.c_int(0),
                              P(h["play"]), P(h["plev"]), P(h["tlay"]), P(h["tlev"]), P(h["tsfc"]),
                              P(h["h2o"]), P(h["o3"]), P(h["co2"]), g("ch4"), g("n2o"), g("o2"),
                              g("cfc11"), g("cfc12"), g("cfc22"), g("ccl4"), NULL,
                              C.c_int(0), C.c_int(0), C.c_int(0), *([NULL] * 6), NULL,
                              P(houts["lw_uflx"]), P(houts["lw_dflx"]), P(houts["lw_hr"]), cs(houts["lw_uflxc"]),
                              cs(houts["lw_dflxc"]), cs(houts["lw_hrc"]), NULL, NULL)
        if rc:
            raise RuntimeError(L_.rrtmg_b200_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        return sharding.max_over_ranks(x, device=dev)

    def timed(fn, steps, warm=1):
        """max over ranks of the CUDA-event time of `steps` calls [ms per step]"""
        for _ in range(warm):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        barrier()
        return max_over_ranks(a.elapsed_time(b)) / steps

    def timed_host(fn, steps):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        return max_over_ranks(1e3 * (time.perf_counter() - t0)) / steps

    # ---------------- device-resident timing (the headline)
    for _ in range(args.warmup):
        step_device()
    L_.rrtmg_b200_set_option(b"kernel_timing", C.c_long(1))
    L_.rrtmg_b200_kernel_times(None, None, C.c_int(1))
    launches0 = rrtmg.launch_count()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    sampler.stop_flag.set()
    launches = rrtmg.launch_count() - launches0
    kms = (C.c_double * len(KERNELS))()
    kn = (C.c_long * len(KERNELS))()
    L_.rrtmg_b200_kernel_times(kms, kn, C.c_int(1))
    L_.rrtmg_b200_set_option(b"kernel_timing", C.c_long(0))
    sampler.join(timeout=2)
    nshort = max(1, min(args.steps, 5))

    # ---------------- the other scaling mode as an extra: weak (one full batch per rank) next to a strong headline
    other = None
    if world > 1:
        if strong:
            W = Dev(batch(seed=20240917 + rank))
            ms_o = timed(lambda: step_device(W), nshort)
            other = {"scaling": "weak", "columns_per_gpu": W.n, "ms_per_step": ms_o, "value": W.n * world / (ms_o * 1e-3),
                     "unit": "columns/s", "what": "every rank one full batch of the named resolution, device-resident"}
            # ... and end to end, as the shim calls it (share_inputs, no clear-sky outputs): all ranks pull on the host memory at once
            hw = {k: torch.empty(v.numel(), dtype=torch.float64).pin_memory() for k, v in W.out.items() if not k.endswith("c")}
            hw.update({k: None for k in W.out if k.endswith("c")})
            L_.rrtmg_b200_set_option(b"share_inputs", C.c_long(1))
            ms_we = timed_host(lambda: step_host(False, W, hw), 3)
            L_.rrtmg_b200_set_option(b"share_inputs", C.c_long(0))
            wl, wv = W.n * nlay, W.n * (nlay + 1)
            other["e2e"] = {"value": W.n * world / (ms_we * 1e-3), "unit": "columns/s", "ms_per_step": ms_we, "steps": 3,
                            "h2d_bytes_per_step": 8 * (5 * wl + 2 * wv + 3 * W.n) if not secondary else None,
                            "d2h_bytes_per_step": 8 * (4 * wv + 2 * wl)}
            del W, hw
        elif nlat_full % world == 0:
            rows = nlat_full // world
            S = Dev(batch((rank * rows, (rank + 1) * rows)))
            ms_o = timed(lambda: step_device(S), nshort)
            other = {"scaling": "strong", "columns_per_gpu": S.n, "ms_per_step": ms_o, "value": nlon * nlat_full / (ms_o * 1e-3),
                     "unit": "columns/s", "what": "one batch of the named resolution, latitude rows split over the ranks"}
            del S
        torch.cuda.empty_cache()

    # ---------------- the same step with SW and LW on two streams (extra, untimed for `value`)
    ms_two = None
    if not two_streams:
        side2 = torch.cuda.Stream(device=dev)
        sh2 = C.c_void_p(side2.cuda_stream)

        def step_two():
            nonlocal sh_sw
            keep = sh_sw
            sh_sw = sh2
            side2.wait_stream(stream)
            try:
                step_device()
            finally:
                sh_sw = keep
            stream.wait_stream(side2)
        ms_two = timed(step_two, nshort)

    # ---------------- the same step with a realistic instantaneous sun (about half the columns at night; SURVEY.md 8d:
    #                  the headline keeps every column sunlit, this second number is reported next to it)
    ms_night = None
    night_frac = None
    if os.environ.get("RRTMG_SKIP_NIGHT") != "1":
        lon = np.arange(nlon) * (2 * np.pi / nlon)
        j0 = rank * (nlat_full // world) if strong else 0
        lat1 = np.arcsin(np.linspace(-1.0 + 1.0 / nlat_full, 1.0 - 1.0 / nlat_full, nlat_full))[j0:j0 + cols.nlat]
        cz = (np.cos(lat1)[None, :] * np.cos(lon - np.pi)[:, None]).ravel(order="F")
        cz = np.where(cz < 0.0, 0.0, cz)
        night_frac = float((cz <= 0.0).mean())
        keep_cz = D.d["coszen"]
        D.d["coszen"] = torch.from_numpy(cz).to(dev)
        ms_night = timed(step_device, nshort)
        D.d["coszen"] = keep_cz

    # ---------------- end to end through the host-pointer ABI
    # (a) as MiMA's shim calls it: the SW call's device copies of the shared inputs serve the LW call (option share_inputs;
    #     run_rrtmg passes the same arrays to both, rrtm_radiation.f90:686-712, 722-748), no clear-sky outputs
    L_.rrtmg_b200_set_option(b"share_inputs", C.c_long(1))
    ms_e2e = timed_host(lambda: step_host(False), nshort)
    L_.rrtmg_b200_set_option(b"share_inputs", C.c_long(0))
    nsec = 7 if secondary else 0
    h2d = 8 * ((5 + 3 * (1 if secondary else 0)) * nl + 2 * nv + 3 * ncol) + 8 * (4 * (1 if secondary else 0)) * nl   # shared inputs once (+ LW-only CFCs)
    if secondary:      # LW-only array inputs present: the library falls back to separate uploads for the LW call
        h2d = 8 * (8 * nl + 2 * nv + 3 * ncol) + 8 * (12 * nl + 2 * nv + 1 * ncol)
    d2h = 8 * (4 * nv + 2 * nl)
    # (b) the plain two calls: every input uploaded by each call, all twelve output arrays copied back
    ms_e2e_all = timed_host(lambda: step_host(True), nshort)
    h2d_all = 8 * ((5 + (3 if secondary else 0)) * nl + 2 * nv + 3 * ncol) + 8 * ((5 + nsec) * nl + 2 * nv + 1 * ncol)
    d2h_all = 2 * 8 * (4 * nv + 2 * nl)

    # ---------------- the whole radiation step of run_rrtmg through the C ABI (device-side marshaling, interp_temp
    #                  and compute_zenith): GCM state in, heating rate + 2-D fields out
    from mima_b200 import rrtm_radiation as rr
    from mima_b200.columns import gcm_state_from_columns
    gs = gcm_state_from_columns(cols)
    gh = {k: pin(gs[k]) for k in ("lat", "lon", "p_full", "p_half", "albedo", "q", "t", "t_surf", "z_full", "z_half", "o3f", "tdt")}
    go = {"coszen": torch.empty(ncol, dtype=torch.float64).pin_memory(), "flux_sw": torch.empty(ncol, dtype=torch.float64).pin_memory(),
          "flux_lw": torch.empty(ncol, dtype=torch.float64).pin_memory(), "tdt_rad": torch.empty(nl, dtype=torch.float64).pin_memory()}
    rcfg = rr.RadConfig(co2ppmv=float(colkw.get("co2_ppmv", 390.0)), solr_cnst=cols.scon).to_c()

    def step_run_rrtmg():
        rc = L_.rrtmg_b200_run_rrtmg(C.byref(rcfg), C.c_int(nlon), C.c_int(cols.nlat), C.c_int(nlay), C.c_int(0), C.c_int(90),
                                     P(gh["lat"]), P(gh["lon"]), P(gh["p_full"]), P(gh["p_half"]), P(gh["albedo"]),
                                     P(gh["q"]), P(gh["t"]), P(gh["t_surf"]), P(gh["z_full"]), P(gh["z_half"]), NULL,
                                     P(gh["o3f"]), P(gh["tdt"]), P(go["coszen"]), P(go["flux_sw"]), P(go["flux_lw"]),
                                     P(go["tdt_rad"]), NULL, NULL, NULL, NULL, NULL)
        if rc:
            raise RuntimeError(L_.rrtmg_b200_last_error().decode())

    ms_rr = timed_host(step_run_rrtmg, nshort)
    rr_h2d = 8 * (6 * nl + 2 * nv + 4 * ncol)          # p_full q t z_full o3f tdt | p_half z_half | lat lon albedo t_surf
    rr_d2h = 8 * (2 * nl + 3 * ncol)                   # tdt tdt_rad | coszen flux_sw flux_lw

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    ms_step = ms_dev / args.steps
    value = total_cols / (ms_step * 1e-3)
    kab = kernel_alg_bytes(nlay)
    prof, prof_src = profiled(args.workload)
    per_kernel = {}
    for i, k in enumerate(KERNELS):
        if kn[i]:
            avg_ms = kms[i] / kn[i]
            cols_per_launch = ncol * args.steps / kn[i]
            per_kernel[k] = {"ms_per_step": kms[i] / args.steps, "launches_per_step": kn[i] / args.steps,
                             "gbs": kab[k] * cols_per_launch / (avg_ms * 1e-3) / 1e9,
                             "frac": kab[k] * cols_per_launch / (avg_ms * 1e-3) / 1e9 / peak,
                             "fp64_pipe_pct": prof[k]["fp64_pipe_pct"] if prof and k in prof else None,
                             "dram_bytes_per_column": prof[k]["dram_bytes_per_column"] if prof and k in prof else None}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"]) if per_kernel else None
    roof = None
    if dom:
        ach = per_kernel[dom]["gbs"]
        cpl = int(ncol * args.steps / kn[KERNELS.index(dom)])
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": prof[dom]["dram_bytes_per_column"] * cpl if prof and dom in prof else None,
                "traffic_source": prof_src,
                "fp64_pipe_pct": per_kernel[dom]["fp64_pipe_pct"],
                "peak_source": peak_src,
                "step_achieved": b_alg(nlay) * ncol / (ms_step * 1e-3) / 1e9,
                "step_frac": b_alg(nlay) * ncol / (ms_step * 1e-3) / 1e9 / peak,
                "kernel_share_of_step": per_kernel[dom]["ms_per_step"] / ms_step,
                "per_kernel": per_kernel}
    cpu = None
    if not args.no_cpu:
        rate, nt, sample, times, kind = cpu_reference_rate(args.workload, args.cpu_sample, 3)
        r1, _, s1, _, _ = cpu_reference_rate(args.workload, max(512, args.cpu_sample // 8), 1, threads=1)
        cpu = {"value": rate, "unit": "columns/s", "cores": nt, "kind": kind,
               "sample": f"{sample} columns x {nlay} layers of {args.workload} (mid-latitude rows), best of 3, threads over column blocks",
               "what": CPU_KINDS[kind], "times_s": times,
               "one_core": {"value": r1, "unit": "columns/s", "sample": f"{s1} columns, one thread"}}

    def e2e(ms, hi, ho, what):
        return {"value": total_cols / (ms * 1e-3), "unit": "columns/s", "ms_per_step": ms, "h2d_bytes_per_step": hi,
                "d2h_bytes_per_step": ho, "steps": nshort, "what": what}

    line = {
        "metric": "RRTMG LW+SW columns/sec", "value": value, "unit": "columns/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.workload, world, strong, 2 if two_streams else 1),
        "e2e": e2e(ms_e2e, h2d, d2h, "rrtmg_b200_sw + rrtmg_b200_lw on pinned host arrays as the Fortran shim calls them: option share_inputs = 1, "
                                     "NULL clear-sky outputs (b200_clear_sky_outputs = .false.)"),
        "e2e_all_outputs": e2e(ms_e2e_all, h2d_all, d2h_all, "the same two calls without share_inputs and with all twelve output arrays"),
        "e2e_run_rrtmg": e2e(ms_rr, rr_h2d, rr_d2h, "rrtmg_b200_run_rrtmg with host buffers: marshaling + interp_temp + compute_zenith (daily-mean sun) + SW + LW"),
        ("weak_scaling" if strong else "strong_scaling"): other,
        "two_streams_ms_per_step": ms_two,
        "realistic_night": None if ms_night is None else {"ms_per_step": ms_night, "night_fraction": night_frac,
                                                           "value": total_cols / (ms_night * 1e-3), "unit": "columns/s"},
        "gpu_launches": int(launches),
        "roofline": roof, "cpu_baseline": cpu, "clocks": sampler.summary(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
