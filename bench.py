#!/usr/bin/env python3
"""bench.py -- RRTMG LW+SW columns/second on B200 (metric of BASELINE.json), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload T170L60] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path -- rrtmg_sw + rrtmg_lw, as run_rrtmg calls them
(rrtm_radiation.f90:686-748) -- over one batch of seeded synthetic columns (SURVEY.md 8d generator).

  value      whole-job columns/s with the inputs resident in HBM (device-pointer ABI, CUDA events on the
             launching stream, max over ranks).
  e2e        the same metric through the host-pointer C ABI (rrtmg_b200_sw / rrtmg_b200_lw): pinned host
             inputs -> H2D -> kernels -> D2H of all six output arrays, every step.
  roofline   dominant kernel: its share of the algorithmic bytes B_alg(L) = 8*(1060 L + 35) B/column
             (per-kernel split in DESIGN.md) divided by its CUDA-event duration, against the
             measured HBM copy bandwidth in MEASURED_PEAKS.json.  `step_frac` is the same for the whole step.
  cpu_baseline  the C oracle (a port of the reference Fortran; the reference itself cannot be compiled
             here) on all host cores, on a bounded column sample of the same workload.

Multi-GPU: columns are independent, each rank owns a block of latitude rows and drives one GPU; no
collective on the data path (the only communication is the timing reduction).  Weak scaling: every rank
processes one full batch of the named resolution (rank r = latitude-row block r of an N-times taller grid).

--impl reference times the CPU implementation (oracle port, OpenMP over columns, all host threads) on the
same workload/metric; rank 0 only.
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

KERNELS = ["lw_prep", "lw_taumol", "lw_rtrn", "sw_prep", "sw_taumol", "sw_solver"]


def kernel_alg_bytes(L: int) -> dict:
    """Split of B_alg(L) = 8*(1060 L + 35) bytes/column over the six kernels."""
    return {
        "lw_prep": 8 * (30 * L + 19),          # LW interface inputs actually dereferenced
        "lw_taumol": 8 * (280 * L),            # write taug, fracs (140 g)
        "lw_rtrn": 8 * (280 * L + 6 * L + 4),  # read them back + LW outputs
        "sw_prep": 8 * (10 * L + 8),
        "sw_taumol": 8 * (224 * L),            # write taug, taur (112 g)
        "sw_solver": 8 * (224 * L + 6 * L + 4),
    }


def b_alg(L: int) -> int:
    return 8 * (1060 * L + 35)


def measured_traffic(workload: str, kernel: str, ncol: int):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of this workload
    (profiles/traffic.json: bytes per column per launch, measured on a full-size pass), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))[workload][kernel]
        return t["dram_bytes_per_column"] * ncol
    except Exception:
        return None


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


def cpu_reference_rate(workload: str, sample_cols: int, repeats: int, threads: int | None = None):
    """columns/s of the CPU implementation (oracle port) on a bounded sample of the workload."""
    from mima_b200.columns import RESOLUTIONS, make_columns
    from oracle.pyoracle import Oracle
    nlon, nlat, nlay = RESOLUTIONS[workload]
    rows = max(1, min(nlat, sample_cols // nlon))
    j0 = (nlat - rows) // 2
    cols = make_columns(workload, lat_rows=(j0, j0 + rows))
    orc = Oracle()
    # all host cores this process may run on (torchrun exports OMP_NUM_THREADS=1, which must not shrink the CPU arm)
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    nt = threads or max(orc.max_threads, ncpu)
    best = None
    times = []
    for _ in range(repeats):
        t = time.perf_counter()
        orc.rrtmg_sw(cols, nthreads=nt)
        orc.rrtmg_lw(cols, nthreads=nt)
        dt = time.perf_counter() - t
        times.append(dt)
        best = dt if best is None else min(best, dt)
    return cols.ncol / best, nt, cols.ncol, times


def run_reference(args, rank):
    if rank != 0:
        return
    times_all = []
    rate = None
    sample = None
    nt = None
    for i in range(args.warmup + args.steps):
        r, nt, sample, times = cpu_reference_rate(args.workload, args.cpu_sample, 1)
        if i >= args.warmup:
            times_all.append(times[0])
    total = sum(times_all)
    rate = sample * len(times_all) / total
    from mima_b200.columns import RESOLUTIONS
    nlon, nlat, nlay = RESOLUTIONS[args.workload]
    line = {
        "impl": "reference", "metric": "RRTMG LW+SW columns/sec", "value": rate, "unit": "columns/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times_all),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "columns_per_gpu": nlon * nlat, "layers": nlay,
                   "note": "CPU arm: each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": rate, "unit": "columns/s", "cores": nt, "kind": "port",
                         "sample": f"{sample} columns x {nlay} layers of {args.workload} (mid-latitude rows), OpenMP over columns"},
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
    ap.add_argument("--split", action="store_true",
                    help="strong scaling: the batch of the named resolution is split by latitude rows over the ranks "
                         "(e.g. T341L80 on 8 GPUs = 65536 columns per GPU); default is weak scaling, one full batch per rank")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
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

    nlon, nlat, nlay = RESOLUTIONS[args.workload]
    # weak scaling: every rank holds one full batch; rank r uses its own seed = its own latitude-row block
    if args.split and world > 1:
        if nlat % world:
            raise SystemExit(f"bench.py --split: {nlat} latitude rows do not divide over {world} ranks")
        rows = nlat // world
        cols = make_columns(args.workload, seed=20240917, lat_rows=(rank * rows, (rank + 1) * rows))
        nlat = rows
    else:
        cols = make_columns(args.workload, seed=20240917 + rank)
    ncol = cols.ncol

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).pin_memory()
        return t

    names_l = ["play", "tlay", "h2o", "o3", "co2"]
    names_v = ["plev", "tlev"]
    host = {k: pin(getattr(cols, k)) for k in names_l + names_v + ["tsfc", "albedo", "coszen"]}
    devt = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    nl, nv = ncol * nlay, ncol * (nlay + 1)
    outs = {k: torch.empty(nv if k.endswith("flx") or k.endswith("flxc") else nl, dtype=torch.float64, device=dev)
            for k in ("lw_uflx", "lw_dflx", "lw_hr", "lw_uflxc", "lw_dflxc", "lw_hrc",
                      "sw_uflx", "sw_dflx", "sw_hr", "sw_uflxc", "sw_dflxc", "sw_hrc")}
    for k in ("lw_hr", "lw_hrc", "sw_hr", "sw_hrc"):
        outs[k] = torch.empty(nl, dtype=torch.float64, device=dev)
    houts = {k: torch.empty(v.numel(), dtype=torch.float64).pin_memory() for k, v in outs.items()}
    torch.cuda.synchronize()

    def P(t):
        return C.c_void_p(t.data_ptr())

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

    def step_device(d=None, n=None):
        d = devt if d is None else d
        n = ncol if n is None else n
        if two_streams:
            side.wait_stream(stream)
        rc = L_.rrtmg_b200_sw_device(C.c_int(n), C.c_int(nlay), C.byref(icld), C.byref(iaer),
                                     P(d["play"]), P(d["plev"]), P(d["tlay"]), P(d["tlev"]), P(d["tsfc"]),
                                     P(d["h2o"]), P(d["o3"]), P(d["co2"]), NULL, NULL, NULL,
                                     P(d["albedo"]), P(d["albedo"]), P(d["albedo"]), P(d["albedo"]), P(d["coszen"]),
                                     C.c_double(cols.adjes), C.c_int(cols.dyofyr), C.c_double(cols.scon),
                                     C.c_int(0), C.c_int(0), C.c_int(0), *([NULL] * 13),
                                     P(outs["sw_uflx"]), P(outs["sw_dflx"]), P(outs["sw_hr"]), P(outs["sw_uflxc"]),
                                     P(outs["sw_dflxc"]), P(outs["sw_hrc"]), sh_sw)
        if rc:
            raise RuntimeError(L_.rrtmg_b200_last_error().decode())
        rc = L_.rrtmg_b200_lw_device(C.c_int(n), C.c_int(nlay), C.byref(icld), C.c_int(0),
                                     P(d["play"]), P(d["plev"]), P(d["tlay"]), P(d["tlev"]), P(d["tsfc"]),
                                     P(d["h2o"]), P(d["o3"]), P(d["co2"]), *([NULL] * 8),
                                     C.c_int(0), C.c_int(0), C.c_int(0), *([NULL] * 6), NULL,
                                     P(outs["lw_uflx"]), P(outs["lw_dflx"]), P(outs["lw_hr"]), P(outs["lw_uflxc"]),
                                     P(outs["lw_dflxc"]), P(outs["lw_hrc"]), NULL, NULL, sh)
        if rc:
            raise RuntimeError(L_.rrtmg_b200_last_error().decode())
        if two_streams:
            stream.wait_stream(side)

    def HP(t):
        return C.c_void_p(t.data_ptr())

    def step_host(clear_sky=True):
        h = host
        cs = (lambda t: HP(t)) if clear_sky else (lambda t: NULL)
        rc = L_.rrtmg_b200_sw(C.c_int(ncol), C.c_int(nlay), C.byref(icld), C.byref(iaer),
                              HP(h["play"]), HP(h["plev"]), HP(h["tlay"]), HP(h["tlev"]), HP(h["tsfc"]),
                              HP(h["h2o"]), HP(h["o3"]), HP(h["co2"]), NULL, NULL, NULL,
                              HP(h["albedo"]), HP(h["albedo"]), HP(h["albedo"]), HP(h["albedo"]), HP(h["coszen"]),
                              C.c_double(cols.adjes), C.c_int(cols.dyofyr), C.c_double(cols.scon),
                              C.c_int(0), C.c_int(0), C.c_int(0), *([NULL] * 13),
                              HP(houts["sw_uflx"]), HP(houts["sw_dflx"]), HP(houts["sw_hr"]), cs(houts["sw_uflxc"]),
                              cs(houts["sw_dflxc"]), cs(houts["sw_hrc"]))
        if rc:
            raise RuntimeError(L_.rrtmg_b200_last_error().decode())
        rc = L_.rrtmg_b200_lw(C.c_int(ncol), C.c_int(nlay), C.byref(icld), C.c_int(0),
                              HP(h["play"]), HP(h["plev"]), HP(h["tlay"]), HP(h["tlev"]), HP(h["tsfc"]),
                              HP(h["h2o"]), HP(h["o3"]), HP(h["co2"]), *([NULL] * 8),
                              C.c_int(0), C.c_int(0), C.c_int(0), *([NULL] * 6), NULL,
                              HP(houts["lw_uflx"]), HP(houts["lw_dflx"]), HP(houts["lw_hr"]), cs(houts["lw_uflxc"]),
                              cs(houts["lw_dflxc"]), cs(houts["lw_hrc"]), NULL, NULL)
        if rc:
            raise RuntimeError(L_.rrtmg_b200_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        return sharding.max_over_ranks(x, device=dev)

    # ---------------- device-resident timing
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
    kms = (C.c_double * 6)()
    kn = (C.c_long * 6)()
    L_.rrtmg_b200_kernel_times(kms, kn, C.c_int(1))
    L_.rrtmg_b200_set_option(b"kernel_timing", C.c_long(0))
    sampler.join(timeout=2)

    # ---------------- strong scaling (extra): the ONE batch of the named resolution cut into `world` blocks of latitude
    #                  rows, every rank timing its block (BASELINE.json: "T170L60 sharded by latitude rows at 1/2/4/8")
    strong = None
    if world > 1 and nlat % world == 0 and not args.split:
        nsub = ncol // world
        dsub = {}
        for k, t in devt.items():
            rows = t.numel() // ncol
            dsub[k] = t.view(rows, ncol)[:, :nsub].contiguous().view(-1)
        for _ in range(args.warmup):
            step_device(dsub, nsub)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(args.steps):
            step_device(dsub, nsub)
        s1.record(stream)
        barrier()
        ms_strong = max_over_ranks(s0.elapsed_time(s1)) / args.steps
        strong = {"columns_total": ncol, "columns_per_gpu": nsub, "ms_per_step": ms_strong,
                  "value": ncol / (ms_strong * 1e-3), "unit": "columns/s",
                  "what": "one batch of the named resolution, latitude rows split over the ranks, device-resident"}
        del dsub

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
        step_two()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        n2 = max(1, min(args.steps, 5))
        for _ in range(n2):
            step_two()
        f1.record(stream)
        barrier()
        ms_two = max_over_ranks(f0.elapsed_time(f1)) / n2

    # ---------------- the same step with a realistic instantaneous sun (about half the columns at night; SURVEY.md 8d:
    #                  the headline keeps every column sunlit, this second number is reported next to it)
    ms_night = None
    night_frac = None
    if os.environ.get("RRTMG_SKIP_NIGHT") != "1":
        lon = np.arange(nlon) * (2 * np.pi / nlon)
        lat1 = np.arcsin(np.linspace(-1.0 + 1.0 / nlat, 1.0 - 1.0 / nlat, nlat))
        cz = (np.cos(lat1)[None, :] * np.cos(lon - np.pi)[:, None]).ravel(order="F")
        cz = np.where(cz < 0.0, 0.0, cz)
        night_frac = float((cz <= 0.0).mean())
        cz_dev = torch.from_numpy(cz).to(dev)
        keep_cz = devt["coszen"]
        devt["coszen"] = cz_dev
        step_device()
        barrier()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record(stream)
        nn = max(1, min(args.steps, 5))
        for _ in range(nn):
            step_device()
        n1.record(stream)
        barrier()
        ms_night = max_over_ranks(n0.elapsed_time(n1)) / nn
        devt["coszen"] = keep_cz

    # ---------------- end-to-end through the host-pointer ABI
    step_host()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        step_host()
    barrier()
    ms_e2e = max_over_ranks(1e3 * (time.perf_counter() - t0))
    h2d = 8 * (5 * nl + 2 * nv + 3 * ncol) + 8 * (5 * nl + 2 * nv + 1 * ncol)   # SW inputs + LW inputs
    d2h = 2 * 8 * (4 * nv + 2 * nl)
    # the same two calls as MiMA's shim makes them: the clear-sky output arrays, which run_rrtmg never reads
    # (rrtm_radiation.f90:716, 752, 790-791), are not requested (NULL), halving the device-to-host volume
    step_host(False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host(False)
    barrier()
    ms_e2e_shim = max_over_ranks(1e3 * (time.perf_counter() - t0))

    # ---------------- the whole radiation step of run_rrtmg through the C ABI (device-side marshaling, interp_temp
    #                  and compute_zenith): GCM state in, heating rate + 2-D fields out
    from mima_b200 import rrtm_radiation as rr
    from mima_b200.columns import gcm_state_from_columns
    gs = gcm_state_from_columns(cols)
    gh = {k: pin(gs[k]) for k in ("lat", "lon", "p_full", "p_half", "albedo", "q", "t", "t_surf", "z_full", "z_half", "o3f", "tdt")}
    go = {"coszen": torch.empty(ncol, dtype=torch.float64).pin_memory(), "flux_sw": torch.empty(ncol, dtype=torch.float64).pin_memory(),
          "flux_lw": torch.empty(ncol, dtype=torch.float64).pin_memory(), "tdt_rad": torch.empty(nl, dtype=torch.float64).pin_memory()}
    rcfg = rr.RadConfig(co2ppmv=390.0, solr_cnst=cols.scon).to_c()

    def step_run_rrtmg():
        rc = L_.rrtmg_b200_run_rrtmg(C.byref(rcfg), C.c_int(nlon), C.c_int(nlat), C.c_int(nlay), C.c_int(0), C.c_int(90),
                                     HP(gh["lat"]), HP(gh["lon"]), HP(gh["p_full"]), HP(gh["p_half"]), HP(gh["albedo"]),
                                     HP(gh["q"]), HP(gh["t"]), HP(gh["t_surf"]), HP(gh["z_full"]), HP(gh["z_half"]), NULL,
                                     HP(gh["o3f"]), HP(gh["tdt"]), HP(go["coszen"]), HP(go["flux_sw"]), HP(go["flux_lw"]),
                                     HP(go["tdt_rad"]), NULL, NULL, NULL, NULL, NULL)
        if rc:
            raise RuntimeError(L_.rrtmg_b200_last_error().decode())

    step_run_rrtmg()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_run_rrtmg()
    barrier()
    ms_rr = max_over_ranks(1e3 * (time.perf_counter() - t0))
    rr_h2d = 8 * (6 * nl + 2 * nv + 4 * ncol)          # p_full q t z_full o3f tdt | p_half z_half | lat lon albedo t_surf
    rr_d2h = 8 * (2 * nl + 3 * ncol)                   # tdt tdt_rad | coszen flux_sw flux_lw

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    ms_step = ms_dev / args.steps
    total_cols = ncol * world
    value = sharding.aggregate_rate(ncol, world, ms_step)
    kab = kernel_alg_bytes(nlay)
    per_kernel = {}
    for i, k in enumerate(KERNELS):
        if kn[i]:
            avg_ms = kms[i] / kn[i]
            cols_per_launch = ncol * args.steps / kn[i]
            per_kernel[k] = {"ms_per_step": kms[i] / args.steps, "launches_per_step": kn[i] / args.steps,
                             "gbs": kab[k] * cols_per_launch / (avg_ms * 1e-3) / 1e9}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"]) if per_kernel else None
    roof = None
    if dom:
        ach = per_kernel[dom]["gbs"]
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": measured_traffic(args.workload, dom, int(ncol * args.steps / kn[KERNELS.index(dom)])),
                "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)",
                "peak_source": peak_src,
                "step_achieved": b_alg(nlay) * ncol / (ms_step * 1e-3) / 1e9,
                "step_frac": b_alg(nlay) * ncol / (ms_step * 1e-3) / 1e9 / peak,
                "kernel_share_of_step": per_kernel[dom]["ms_per_step"] / ms_step,
                "per_kernel": per_kernel}
    cpu = None
    if not args.no_cpu:
        rate, nt, sample, times = cpu_reference_rate(args.workload, args.cpu_sample, 3)
        r1, _, s1, _ = cpu_reference_rate(args.workload, max(512, args.cpu_sample // 8), 1, threads=1)
        cpu = {"value": rate, "unit": "columns/s", "cores": nt, "kind": "port",
               "sample": f"{sample} columns x {nlay} layers of {args.workload} (mid-latitude rows), best of 3, OpenMP over columns",
               "times_s": times, "one_core": {"value": r1, "unit": "columns/s", "sample": f"{s1} columns, one thread"}}
    line = {
        "metric": "RRTMG LW+SW columns/sec", "value": value, "unit": "columns/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if (args.split and world > 1) else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "columns_per_gpu": ncol, "layers": nlay, "grid": f"{nlon}x{nlat}",
                   "sharding": "latitude-row blocks, one rank per GPU, no collective",
                   "l2": "inputs + staging per step exceed the 126 MB L2 (no flush needed)",
                   "streams": 2 if two_streams else 1, "all_sunlit": True, "lw_tables": "synthetic (reference LW k_g file stripped)", "sw_tables": "reference"},
        "e2e": {"value": total_cols / (ms_e2e / e2e_steps * 1e-3), "unit": "columns/s", "ms_per_step": ms_e2e / e2e_steps,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
        "e2e_total_sky_only": {"value": total_cols / (ms_e2e_shim / e2e_steps * 1e-3), "unit": "columns/s", "ms_per_step": ms_e2e_shim / e2e_steps,
                               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h // 2, "steps": e2e_steps,
                               "what": "rrtmg_b200_sw + rrtmg_b200_lw with NULL for the clear-sky outputs MiMA discards (shim switch b200_clear_sky_outputs = .false.)"},
        "e2e_run_rrtmg": {"value": total_cols / (ms_rr / e2e_steps * 1e-3), "unit": "columns/s", "ms_per_step": ms_rr / e2e_steps,
                          "h2d_bytes_per_step": rr_h2d, "d2h_bytes_per_step": rr_d2h, "steps": e2e_steps,
                          "what": "rrtmg_b200_run_rrtmg with host buffers: marshaling + interp_temp + compute_zenith (daily-mean sun) + SW + LW"},
        "strong_scaling": strong,
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
