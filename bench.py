#!/usr/bin/env python
"""bench.py -- throughput of the B200 stencil hot path, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): Gcell-updates/s and fraction of the HBM roofline.

Headline workload, THE SAME AT EVERY N (so value_N / (N * value_1) is a real weak-scaling efficiency):
BASELINE.json configs[4], jacobi_2d on a weak-scaled grid of (N x 10240) x 81920 fp64 cells (13.4 GB of
A + B per GPU, far larger than L2), TSTEPS = 21.  One *step* = one call kernel(TSTEPS, A, B) = 40 sweeps.
  N = 1   through the public C-ABI entry npb_jacobi2d_f64 on arrays resident in HBM, CUDA events on the
          launch stream around every step.
  N > 1   row slabs, ghost depth 7, halo exchange over NCCL (ncclSend/Recv) overlapped with interior
          compute (npbench_b200/distributed.py); barrier + synchronize around every step, max over ranks.
The line also carries
  roofline      dominant kernel vs the measured HBM peak (MEASURED_PEAKS.json): algorithmic AND
                DRAM-counter fraction (`frac`, `frac_dram`), traffic from the committed ncu capture,
  e2e           the same call on pinned HOST arrays (H2D + kernels (+ halos) + D2H inside the timed region),
  parity        one band of rows per rank compared bit for bit with the CPU oracle, outside the timed region,
  cpu_baseline  (N = 1) the reference's NumPy function on a bounded band of the same grid (1 core: NumPy
                runs these ufunc loops single threaded), plus the C/OpenMP port on all cores as `cpu_port`,
  records       (N = 1) the named single-GPU configs of BASELINE.json -- jacobi_2d S (configs[0]),
                heat_3d L (configs[1]), hdiff paper (configs[2]), vadv paper (configs[3]) -- each with its
                own device-timed value, roofline {frac, traffic}, e2e and the wall time of the UNMODIFIED
                NPBench harness (`-f b200` and `-f numpy`, timeit median from npbench.db),
  suite         every kernel x NPBench preset (N = 1) / the other sharded kernels (N > 1),
  clocks        SM clock / throttle reasons sampled via NVML during the timed region.
--impl reference   the CPU arm on the same config/metric: the reference's own NumPy kernel
        (npbench/benchmarks/polybench/jacobi_2d/jacobi_2d_numpy.py:4-10, imported from the staged checkout)
        on a bounded row band of the grid per step; rank 0 only.  Falls back to the oracle port when no
        checkout is present (kind "port").
"""
import argparse
import ctypes
import json
import os
import sqlite3
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gcell-updates/s (fp64 stencil cell updates per second)"
UNIT = "Gcell/s"
WEAK_ROWS, WEAK_COLS, WEAK_TSTEPS = 10240, 81920, 21
WEAK_SWEEPS = 2 * (WEAK_TSTEPS - 1)
DATA = "synthetic (NPBench initialize closed form jacobi_2d.py:6-10, generated on device per slab)"


def workload_label(world):
    """Identical in both arms (the driver compares config.workload)."""
    return "jacobi_2d weak-scaled, (%d x %d) x %d fp64 grid = %d x %d, TSTEPS=%d (%d sweeps per step)" % (
        world, WEAK_ROWS, WEAK_COLS, world * WEAK_ROWS, WEAK_COLS, WEAK_TSTEPS, WEAK_SWEEPS)


def weak_units(world):
    return WEAK_SWEEPS * (world * WEAK_ROWS - 2) * (WEAK_COLS - 2)


def measured_peak():
    """HBM roofline denominator: MEASURED_PEAKS.json (driver-written) else the recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_table():
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ------------------------------------------------------------------ clocks (NVML)
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting",
               0x10: "sync_boost"}

    def __init__(self, device_index, period=0.005):
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self.period, self._stop, self._t, self.ok = period, threading.Event(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = device_index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[device_index])
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:                                   # pragma: no cover
            self.err = repr(e)

    def _sample(self):
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(self.period)

    def __enter__(self):
        if self.ok:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        if self.ok:
            self._sample()            # at least one sample taken while the last timed step is still hot
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None}


# ------------------------------------------------------------------ helpers on the library
def dev_timer(L, fn, flush=True):
    ms = ctypes.c_float()
    if flush:
        L.l2_flush()
    L.timer_start()
    fn()
    L.timer_stop(ctypes.byref(ms))
    return ms.value


def pinned_array(L, shape):
    n = int(np.prod(shape)) * 8
    p = ctypes.c_void_p()
    L.host_alloc(n, ctypes.byref(p))
    buf = (ctypes.c_double * (n // 8)).from_address(p.value)
    a = np.frombuffer(buf, dtype=np.float64).reshape(shape)
    return a, p


def units_of(bench, p):
    if bench == "jacobi_2d":
        return 2 * (p["TSTEPS"] - 1) * (p["NI"] - 2) * (p["NJ"] - 2), 16.0
    if bench == "heat_3d":
        return 2 * (p["TSTEPS"] - 1) * (p["N"] - 2) ** 3, 16.0
    if bench == "fdtd_2d":
        return p["TMAX"] * p["NX"] * p["NY"], 48.0
    if bench == "jacobi_1d":      # widening row: one interior cell written by one sweep, read 8 + write 8
        return 2 * (p["TSTEPS"] - 1) * (p["N"] - 2), 16.0
    if bench == "seidel_2d":      # one interior cell updated by one Gauss-Seidel sweep (in place: read 8 + write 8)
        return (p["TSTEPS"] - 1) * (p["N"] - 2) ** 2, 16.0
    if bench == "channel_flow":   # as cavity_flow, periodic in x: (ny-2) rows x nx columns per pass; steps = the returned count
        return p.get("steps", 1) * (p["nit"] + 2) * p["nx"] * (p["ny"] - 2), 16.0
    if bench == "cavity_flow":    # one interior cell updated by one pass: per time step nit pressure iterations + b + (u, v)
        return p["nt"] * (p["nit"] + 2) * (p["nx"] - 2) * (p["ny"] - 2), 16.0
    if bench == "adi":            # one interior cell solved by one directional sweep (two sweeps per time step)
        return 2 * p["TSTEPS"] * (p["N"] - 2) ** 2, 16.0
    if bench == "hdiff":
        I, J, K = p["I"], p["J"], p["K"]
        return I * J * K, 8.0 * ((I + 4) * (J + 4) + 2 * I * J) / (I * J)
    I, J, K = p["I"], p["J"], p["K"]
    return I * J * K, 8.0 * (6 * I + 1) / I


SUITE = [
    ("jacobi_2d", "S", dict(TSTEPS=50, NI=150, NJ=150)), ("jacobi_2d", "M", dict(TSTEPS=80, NI=350, NJ=350)),
    ("jacobi_2d", "L", dict(TSTEPS=200, NI=700, NJ=700)), ("jacobi_2d", "paper", dict(TSTEPS=1000, NI=2800, NJ=2800)),
    ("jacobi_2d", "scaled-1gpu", dict(TSTEPS=WEAK_TSTEPS, NI=WEAK_ROWS, NJ=WEAK_COLS)),
    ("heat_3d", "S", dict(TSTEPS=25, N=25)), ("heat_3d", "M", dict(TSTEPS=50, N=40)),
    ("heat_3d", "L", dict(TSTEPS=100, N=70)), ("heat_3d", "paper", dict(TSTEPS=500, N=120)),
    ("heat_3d", "scaled-1gpu", dict(TSTEPS=6, N=1024)),
    ("fdtd_2d", "S", dict(TMAX=20, NX=200, NY=220)), ("fdtd_2d", "M", dict(TMAX=60, NX=400, NY=450)),
    ("fdtd_2d", "L", dict(TMAX=150, NX=800, NY=900)), ("fdtd_2d", "paper", dict(TMAX=500, NX=1000, NY=1200)),
    ("fdtd_2d", "scaled-1gpu", dict(TMAX=10, NX=8192, NY=65536)),
    ("hdiff", "S", dict(I=64, J=64, K=60)), ("hdiff", "M", dict(I=128, J=128, K=160)),
    ("hdiff", "L", dict(I=384, J=384, K=160)), ("hdiff", "paper", dict(I=256, J=256, K=160)),
    ("vadv", "S", dict(I=60, J=60, K=40)), ("vadv", "M", dict(I=112, J=112, K=80)),
    ("vadv", "L", dict(I=180, J=180, K=160)), ("vadv", "paper", dict(I=256, J=256, K=160)),
    # widening row
    ("jacobi_1d", "S", dict(TSTEPS=800, N=3200)), ("jacobi_1d", "M", dict(TSTEPS=3000, N=12000)),
    ("jacobi_1d", "L", dict(TSTEPS=8500, N=34000)), ("jacobi_1d", "paper", dict(TSTEPS=4000, N=32000)),
    ("seidel_2d", "S", dict(TSTEPS=8, N=50)), ("seidel_2d", "M", dict(TSTEPS=15, N=100)),
    ("seidel_2d", "L", dict(TSTEPS=40, N=200)), ("seidel_2d", "paper", dict(TSTEPS=100, N=400)),
    # widening row rank 3
    ("channel_flow", "S", dict(ny=61, nx=61, nit=5)), ("channel_flow", "M", dict(ny=121, nx=121, nit=10)),
    ("channel_flow", "L", dict(ny=201, nx=201, nit=20)), ("channel_flow", "paper", dict(ny=101, nx=101, nit=50)),
    ("cavity_flow", "S", dict(ny=61, nx=61, nt=25, nit=5)), ("cavity_flow", "M", dict(ny=121, nx=121, nt=50, nit=10)),
    ("cavity_flow", "L", dict(ny=201, nx=201, nt=100, nit=20)), ("cavity_flow", "paper", dict(ny=101, nx=101, nt=700, nit=50)),
    # widening row rank 2
    ("adi", "S", dict(TSTEPS=5, N=100)), ("adi", "M", dict(TSTEPS=20, N=200)),
    ("adi", "L", dict(TSTEPS=50, N=500)), ("adi", "paper", dict(TSTEPS=100, N=200)),
]


def make_device_case(nb, bench, p, rng):
    """Allocate + initialise device arrays (NPBench initialisers) and return a callable step."""
    L = nb.lib()
    if bench == "jacobi_2d":
        A, B = nb.DeviceArray((p["NI"], p["NJ"])), nb.DeviceArray((p["NI"], p["NJ"]))
        L.init_jacobi2d_f64(p["NJ"], 0, p["NI"], p["NJ"], A.ptr, B.ptr)
        return (A, B), (lambda: nb.jacobi_2d(p["TSTEPS"], A, B))
    if bench == "heat_3d":
        n = p["N"]
        A, B = nb.DeviceArray((n, n, n)), nb.DeviceArray((n, n, n))
        L.init_heat3d_f64(n, 0, n, A.ptr, B.ptr)
        return (A, B), (lambda: nb.heat_3d(p["TSTEPS"], A, B))
    if bench == "fdtd_2d":
        a = [nb.DeviceArray((p["NX"], p["NY"])) for _ in range(3)] + [nb.DeviceArray((p["TMAX"],))]
        L.init_fdtd2d_f64(p["TMAX"], p["NX"], p["NY"], 0, p["NX"], *(x.ptr for x in a))
        return a, (lambda: nb.fdtd_2d(p["TMAX"], *a))
    if bench == "jacobi_1d":      # jacobi_1d.py:6-10
        n = p["N"]
        A = nb.DeviceArray.from_host((np.arange(n, dtype=np.float64) + 2.0) / n)
        B = nb.DeviceArray.from_host((np.arange(n, dtype=np.float64) + 3.0) / n)
        return (A, B), (lambda: nb.jacobi_1d(p["TSTEPS"], A, B))
    if bench == "seidel_2d":      # seidel_2d.py:6-10
        n = p["N"]
        i, j = np.meshgrid(np.arange(n, dtype=np.float64), np.arange(n, dtype=np.float64), indexing="ij")
        A = nb.DeviceArray.from_host((i * (j + 2.0) + 2.0) / n)
        return (A,), (lambda: nb.seidel_2d(p["TSTEPS"], n, A))
    if bench == "channel_flow":   # channel_flow.py: u = v = 0, p = 1
        nx, ny = p["nx"], p["ny"]
        f0 = [nb.DeviceArray.from_host(a) for a in (np.zeros((ny, nx)), np.zeros((ny, nx)), np.ones((ny, nx)))]
        f = [nb.DeviceArray((ny, nx)) for _ in range(3)]
        dx, dy, dt = 2 / (nx - 1), 2 / (ny - 1), .1 / ((nx - 1) * (ny - 1))

        def step():     # restart from the initial fields; the call blocks until the flow has converged
            for x, x0 in zip(f, f0):
                L.d2d(x.ptr, x0.ptr, nx * ny * 8)
            p["steps"] = nb.channel_flow(p["nit"], f[0], f[1], dt, dx, dy, f[2], 1.0, 0.1, 1.0)
        return (f, f0), step
    if bench == "cavity_flow":    # cavity_flow.py:6-13: zero fields, dx = 2/(nx-1), dy = 2/(ny-1), dt = .1/((nx-1)(ny-1))
        nx, ny = p["nx"], p["ny"]
        z = np.zeros((ny, nx))
        f0 = [nb.DeviceArray.from_host(z) for _ in range(3)]
        f = [nb.DeviceArray((ny, nx)) for _ in range(3)]
        dx, dy, dt = 2 / (nx - 1), 2 / (ny - 1), .1 / ((nx - 1) * (ny - 1))

        def step():     # restart from the initial fields like the harness does (copy_func in setup_str)
            for x, x0 in zip(f, f0):
                L.d2d(x.ptr, x0.ptr, nx * ny * 8)
            nb.cavity_flow(nx, ny, p["nt"], p["nit"], f[0], f[1], dt, dx, dy, f[2], 1.0, 0.1)
        return (f, f0), step
    if bench == "adi":            # adi.py: u = (i + N - j) / N
        n = p["N"]
        i, j = np.meshgrid(np.arange(n, dtype=np.float64), np.arange(n, dtype=np.float64), indexing="ij")
        u0 = nb.DeviceArray.from_host((i + n - j) / n)
        u = nb.DeviceArray((n, n))

        def step():     # the reference's coefficients make u grow by orders of magnitude per call: restart from u0 every time
            L.d2d(u.ptr, u0.ptr, n * n * 8)
            nb.adi(p["TSTEPS"], n, u)
        return (u, u0), step
    I, J, K = p["I"], p["J"], p["K"]
    if bench == "hdiff":   # hdiff.py:6-15 draws U[0,1); any U[0,1) data has the same cost
        a = [nb.DeviceArray.from_host(rng.random(s)) for s in ((I + 4, J + 4, K), (I, J, K), (I, J, K))]
        return a, (lambda: nb.hdiff(*a))
    a = [nb.DeviceArray.from_host(rng.random(s)) for s in
         ((I, J, K), (I, J, K), (I + 1, J, K), (I, J, K), (I, J, K))]
    return a, (lambda: nb.vadv(*a, 0.15))


def run_suite(nb, peak):
    L = nb.lib()
    rng = np.random.default_rng(42)
    rows = []
    for bench, preset, p in SUITE:
        try:
            keep, step = make_device_case(nb, bench, p, rng)
            n0 = L.launch_count()
            step(); L.sync()
            units, bpu = units_of(bench, p)
            launches = int(L.launch_count() - n0)
            est = dev_timer(L, step)
            reps = 3 if est > 20 else (5 if est > 2 else 15)
            ts = [dev_timer(L, step) for _ in range(reps)]
            ms = float(np.median(ts))
            gc = units / (ms * 1e-3) / 1e9
            rows.append({"kernel": bench, "preset": preset, "ms": round(ms, 4), "value": round(gc, 2),
                         "GBps_algorithmic": round(gc * bpu, 1), "frac_of_peak": round(gc * bpu / peak, 3),
                         "launches": launches})
            del keep, step
            L.pool_trim()
        except Exception as e:                                   # keep the headline alive
            rows.append({"kernel": bench, "preset": preset, "error": str(e)[:200]})
    return rows


# ------------------------------------------------------------------ reference checkout (NumPy arm)
def find_checkout():
    try:
        from npbench_b200 import overlay
        return overlay.find_reference()
    except Exception:
        return None


def numpy_kernel(bench_rel, module, func):
    """Import a reference NumPy kernel from the checkout (None when no checkout is present)."""
    ref = find_checkout()
    if ref is None:
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import importlib
    return getattr(importlib.import_module("npbench.benchmarks.%s.%s" % (bench_rel.replace("/", "."), module)), func)


def jacobi_band_numpy(rows):
    """Rows [0, rows) of the global grid with jacobi_2d.py:7-8's closed form (N = WEAK_COLS)."""
    n = WEAK_COLS
    A = np.fromfunction(lambda i, j: i * (j + 2) / n, (rows, n), dtype=np.float64)
    B = np.fromfunction(lambda i, j: i * (j + 3) / n, (rows, n), dtype=np.float64)
    return A, B


def numpy_band_rate(target_s):
    """Pick a band height whose NumPy step takes about target_s; returns (kernel, rows)."""
    kern = numpy_kernel("polybench/jacobi_2d", "jacobi_2d_numpy", "kernel")
    A, B = jacobi_band_numpy(34)                # 32 interior rows x 81920 columns: already out of cache
    t0 = time.perf_counter(); kern(WEAK_TSTEPS, A, B); est = time.perf_counter() - t0
    rows = 2 + int(max(8, min(WEAK_ROWS - 2, 32 * target_s / max(est, 1e-6))))
    return kern, rows


def cpu_baseline_numpy(budget_s=12.0):
    """cpu_baseline of the b200 arm: the reference NumPy kernel on a bounded band, one run."""
    kern, rows = numpy_band_rate(budget_s)
    A, B = jacobi_band_numpy(rows)
    t0 = time.perf_counter(); kern(WEAK_TSTEPS, A, B); t = time.perf_counter() - t0
    units = WEAK_SWEEPS * (rows - 2) * (WEAK_COLS - 2)
    return {"value": round(units / t / 1e9, 4), "unit": UNIT, "cores": 1, "kind": "reference",
            "host_cores": os.cpu_count(),
            "sample": "reference NumPy kernel (jacobi_2d_numpy.py:4-10) on a %d-row band x %d columns of the grid, "
                      "TSTEPS=%d, one run of %.1f s; NumPy evaluates these ufunc loops on 1 core" % (
                          rows, WEAK_COLS, WEAK_TSTEPS, t)}


def cpu_port(budget_s=6.0):
    import oracle
    threads = oracle.max_threads()
    oracle.set_threads(threads)
    rows = 2048
    A, B = oracle.init_jacobi_2d(WEAK_COLS, row0=0, nrows=rows, ncols=WEAK_COLS)
    oracle.jacobi_2d(WEAK_TSTEPS, A, B)
    times, t_end = [], time.perf_counter() + budget_s
    while len(times) < 2 or (time.perf_counter() < t_end and len(times) < 20):
        t0 = time.perf_counter(); oracle.jacobi_2d(WEAK_TSTEPS, A, B); times.append(time.perf_counter() - t0)
    t = float(np.median(times))
    units = WEAK_SWEEPS * (rows - 2) * (WEAK_COLS - 2)
    return {"value": round(units / t / 1e9, 4), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "C/OpenMP oracle port (oracle/stencil_oracle.c, NumPy evaluation order) on a %d-row band, "
                      "median of %d runs" % (rows, len(times))}


def reference_arm(args, world, rank):
    """--impl reference: the reference's NumPy kernel on host cores, same metric/config; rank 0 only."""
    if rank != 0:
        return
    kern = numpy_kernel("polybench/jacobi_2d", "jacobi_2d_numpy", "kernel")
    if kern is not None:
        _, rows = numpy_band_rate(1.5)            # a few seconds per step keeps K + W steps within a few minutes
        A, B = jacobi_band_numpy(rows)
        step = lambda: kern(WEAK_TSTEPS, A, B)
        kind, cores = "reference", 1
        sample = ("reference NumPy kernel (jacobi_2d_numpy.py:4-10) per step on a %d-row band x %d columns of the "
                  "%d-row grid (same columns, same TSTEPS); NumPy evaluates these ufunc loops on 1 core "
                  "(host has %d)" % (rows, WEAK_COLS, world * WEAK_ROWS, os.cpu_count()))
    else:
        import oracle
        cores = max(oracle.max_threads(), os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: override
        oracle.set_threads(cores)
        rows = 2048
        A, B = oracle.init_jacobi_2d(WEAK_COLS, row0=0, nrows=rows, ncols=WEAK_COLS)
        step = lambda: oracle.jacobi_2d(WEAK_TSTEPS, A, B)
        kind = "port"
        sample = "no NPBench checkout on this box: C/OpenMP oracle port on a %d-row band per step, %d threads" % (rows, cores)
    units = WEAK_SWEEPS * (rows - 2) * (WEAK_COLS - 2)
    for _ in range(args.warmup):
        step()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter(); step(); times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    val = units / t / 1e9
    line = {"metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(t * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (NPBench initialize closed form jacobi_2d.py:6-10)",
            "impl": "reference", "config": {"workload": workload_label(world)},
            "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ the unmodified harness (wall time)
def harness_leg(bench, preset, r_b200=10, r_numpy=3, timeout=600):
    """Run the reference's own CLI (run_benchmark.py:49-57 -> Test.run, test.py:53-163) with `-f b200`
    and with `-f numpy`; return the timeit medians it stored in npbench.db (utilities.py:135-151)."""
    ref = find_checkout()
    if ref is None:
        return {"note": "no NPBench checkout on this box (baseline/_ref not staged)"}
    out = {"cli": "run_benchmark.py -b %s -f <fw> -p %s" % (bench, preset)}
    tmp = tempfile.mkdtemp(prefix="npb_harness_")
    env = dict(os.environ, PYTHONPATH=ROOT)
    env.pop("OMP_NUM_THREADS", None)
    for fw, r in (("b200", r_b200), ("numpy", r_numpy)):
        try:
            p = subprocess.run([sys.executable, "-m", "npbench_b200.run", "--reference", ref, "--overlay",
                                os.path.join(tmp, "ov"), "--", "-b", bench, "-f", fw, "-p", preset, "-r", str(r)],
                               cwd=tmp, env=env, capture_output=True, text=True, timeout=timeout)
            con = sqlite3.connect(os.path.join(tmp, "npbench.db"))
            rows = con.execute("SELECT validated, time FROM results WHERE framework = ? AND preset = ?", (fw, preset)).fetchall()
            con.close()
            if p.returncode != 0 or not rows:
                out[fw] = {"error": (p.stderr or p.stdout)[-300:]}
                continue
            out[fw] = {"wall_ms_median": round(float(np.median([t for _, t in rows])) * 1e3, 4), "repeat": len(rows)}
            if fw == "b200":
                out[fw]["validated"] = int(all(v == 1 for v, _ in rows))
                out[fw]["validation_line"] = "validation: SUCCESS" in p.stdout
        except Exception as e:
            out[fw] = {"error": repr(e)[:300]}
    return out


# ------------------------------------------------------------------ named single-GPU records
RECORDS = [   # (name, bench, preset, params, BASELINE.json config, dominant kernel, NumPy arg builder)
    ("jacobi_2d_S", "jacobi_2d", "S", dict(TSTEPS=50, NI=150, NJ=150), "configs[0]"),
    ("heat_3d_L", "heat_3d", "L", dict(TSTEPS=100, N=70), "configs[1]"),
    ("hdiff_paper", "hdiff", "paper", dict(I=256, J=256, K=160), "configs[2]"),
    ("vadv_paper", "vadv", "paper", dict(I=256, J=256, K=160), "configs[3]"),
]


def record_kernel_name(L, bench):
    if bench == "heat_3d":
        return {1: "heat3d_resident_kernel", 2: "heat3d_sweep_kernel", 5: "heat3d_march_kernel",
                6: "heat3d_regtile_kernel"}.get(int(L.heat3d_last_path()), "heat3d")
    if bench == "jacobi_2d":
        return {1: "jacobi2d_regtile_kernel", 2: "jacobi2d_block_kernel", 3: "jacobi2d_march_kernel"}.get(
            int(L.jacobi2d_last_path()), "jacobi2d")
    if bench == "hdiff":
        return {1: "hdiff_march_kernel", 2: "hdiff_ring_kernel"}.get(int(L.hdiff_last_path()), "hdiff")
    return {1: "vadv_pipeline_kernel", 2: "vadv_tma_kernel", 3: "vadv_stream_kernel (TMEM + TMA streaming Thomas solver)"}.get(
        int(L.vadv_last_path()), "vadv")


def host_case(nb, bench, p, rng):
    """Pinned host arrays with NPBench's inputs + the public host-buffer call (the e2e leg)."""
    import oracle
    L = nb.lib()
    if bench == "jacobi_2d":
        src = oracle.init_jacobi_2d(p["NI"])
        call = lambda a: nb.jacobi_2d(p["TSTEPS"], a[0], a[1])
        outs = 2
    elif bench == "heat_3d":
        src = oracle.init_heat_3d(p["N"])
        call = lambda a: nb.heat_3d(p["TSTEPS"], a[0], a[1])
        outs = 2
    elif bench == "hdiff":
        src = oracle.init_hdiff(p["I"], p["J"], p["K"])
        call = lambda a: nb.hdiff(a[0], a[1], a[2])
        outs = 1
    else:
        dtr, *src = oracle.init_vadv(p["I"], p["J"], p["K"])
        call = lambda a: nb.vadv(a[0], a[1], a[2], a[3], a[4], dtr)
        outs = 1
    arrs, ptrs = [], []
    for s in src:
        a, ptr = pinned_array(L, s.shape)
        a[...] = s
        arrs.append(a); ptrs.append(ptr)
    h2d = sum(a.nbytes for a in arrs)
    d2h = sum(a.nbytes for a in arrs[:outs]) if bench != "hdiff" else arrs[1].nbytes
    return arrs, ptrs, call, h2d, d2h


def run_records(nb, peak, peak_src, steps, with_harness):
    L = nb.lib()
    rng = np.random.default_rng(42)
    traffic = traffic_table()
    out = {}
    for name, bench, preset, p, cfg in RECORDS:
        try:
            keep, step = make_device_case(nb, bench, p, rng)
            units, bpu = units_of(bench, p)
            for _ in range(5):
                step()
            L.sync()
            n0 = L.launch_count()
            ts = [dev_timer(L, step, flush=True) for _ in range(steps)]
            launches = int(L.launch_count() - n0) // steps
            ms = float(np.mean(ts))
            val = units / (ms * 1e-3) / 1e9
            alg = bpu * units / max(1, launches)
            us = ms * 1e3 / max(1, launches)
            ach = alg / (us * 1e-6) / 1e9
            tr = traffic.get(name + "_bytes_per_launch")
            rec = {"config": cfg, "workload": "%s preset %s %s, fp64" % (bench, preset, json.dumps(p).replace('"', "")),
                   "value": round(val, 3), "unit": UNIT, "ms_per_step": round(ms, 5), "ms_min": round(min(ts), 5), "steps": steps,
                   "l2": "flushed before every timed step",
                   "roofline": {"bound": "hbm", "kernel": record_kernel_name(L, bench), "achieved": round(ach, 1), "peak": peak,
                                "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": tr,
                                "traffic_source": traffic.get(name + "_source"), "peak_source": peak_src,
                                "algorithmic_bytes_per_launch": alg, "launches_per_step": launches,
                                "avg_launch_us": round(us, 3)}}
            # Single-pass records whose footprint is at least twice the L2 (hdiff / vadv `paper`: 254 / 504 MB against
            # 126 MB) also get the other timing the contract allows for inputs larger than L2: ten launches between ONE
            # pair of events, no flush.  The ~4 us between an event and a lone kernel (3 - 8 % of these 50 - 110 us launches)
            # amortise; `frac` above stays the conservative flushed single-launch figure.
            if launches == 1 and bench in ("hdiff", "vadv") and alg >= 2.0 * 126e6:   # single-pass kernels: algorithmic bytes = footprint
                reps = 10
                L.sync()
                msb = ctypes.c_float()
                L.timer_start()
                for _ in range(reps):
                    step()
                L.timer_stop(ctypes.byref(msb))
                usb = msb.value * 1e3 / reps
                rec["roofline"]["back_to_back"] = {
                    "avg_launch_us": round(usb, 3), "frac": round(alg / (usb * 1e-6) / 1e9 / peak, 4), "launches": reps,
                    "note": "ten launches between one pair of CUDA events, no L2 flush (inputs %.1fx the L2)" % (alg / 126e6)}
            del keep, step
            # e2e: pinned host arrays through the public host-buffer API (H2D + kernels + D2H per call)
            arrs, ptrs, call, h2d, d2h = host_case(nb, bench, p, rng)
            for _ in range(3):
                call(arrs)
            tt = []
            for _ in range(max(5, min(steps, 20))):
                L.l2_flush(); L.sync()
                t0 = time.perf_counter(); call(arrs); tt.append(time.perf_counter() - t0)
            e = float(np.mean(tt))
            rec["e2e"] = {"value": round(units / e / 1e9, 3), "unit": UNIT, "ms_per_step": round(e * 1e3, 4),
                          "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                          "api": "npbench_b200.%s(...) on pinned host ndarrays -> npb_*_f64_host" % bench}
            for ptr in ptrs:
                L.host_free(ptr)
            del arrs
            L.pool_trim()
            if with_harness:
                h = harness_leg(bench, preset)
                if "b200" in h and "wall_ms_median" in h["b200"]:
                    h["b200"]["value"] = round(units / (h["b200"]["wall_ms_median"] * 1e-3) / 1e9, 3)
                if "numpy" in h and "wall_ms_median" in h["numpy"]:
                    h["numpy"]["value"] = round(units / (h["numpy"]["wall_ms_median"] * 1e-3) / 1e9, 4)
                    h["numpy"]["cores"] = 1
                rec["harness"] = h
            out[name] = rec
        except Exception as e:                                   # keep the headline alive
            out[name] = {"error": repr(e)[:300]}
    return out


# ------------------------------------------------------------------ parity of the headline (outside the timed region)
PARITY_BAND = 16       # rows compared per band
def band_parity(fetch_rows, n_rows_global, centre):
    """Compare global rows [centre - 8, centre + 8) after ONE step from the initial state with the CPU oracle
    run on those rows plus the WEAK_SWEEPS-row dependency cone on either side.  fetch_rows(lo, hi) returns the
    (A, B) rows [lo, hi) of the device result as NumPy arrays."""
    import oracle
    oracle.set_threads(max(oracle.max_threads(), os.cpu_count() or 1))
    lo = max(0, centre - PARITY_BAND // 2); hi = min(n_rows_global, lo + PARITY_BAND)
    b_lo = max(0, lo - WEAK_SWEEPS); b_hi = min(n_rows_global, hi + WEAK_SWEEPS)
    A, B = oracle.init_jacobi_2d(WEAK_COLS, row0=b_lo, nrows=b_hi - b_lo, ncols=WEAK_COLS)
    if b_hi < n_rows_global or b_lo > 0:
        pass   # band edges act as fixed borders; their error front moves one row per sweep and stays outside [lo, hi)
    oracle.jacobi_2d(WEAK_TSTEPS, A, B)
    gA, gB = fetch_rows(lo, hi)
    ok = bool(np.array_equal(gA, A[lo - b_lo:hi - b_lo]) and np.array_equal(gB, B[lo - b_lo:hi - b_lo]))
    return {"rows": [lo, hi], "bit_exact": ok}


# ------------------------------------------------------------------ N = 1
def single_gpu(args):
    import npbench_b200 as nb
    nb.init(0)
    L = nb.lib()
    peak, peak_src = measured_peak()
    traffic = traffic_table()
    units = weak_units(1)
    A, B = nb.DeviceArray((WEAK_ROWS, WEAK_COLS)), nb.DeviceArray((WEAK_ROWS, WEAK_COLS))
    init = lambda: L.init_jacobi2d_f64(WEAK_COLS, 0, WEAK_ROWS, WEAK_COLS, A.ptr, B.ptr)
    step = lambda: nb.jacobi_2d(WEAK_TSTEPS, A, B)
    init()
    warm = max(3, args.warmup)
    for _ in range(warm):
        step()
    L.sync()
    times = []
    n0 = L.launch_count()
    with ClockSampler(0) as clk:
        for _ in range(args.steps):
            times.append(dev_timer(L, step, flush=False))        # 13.4 GB of inputs: nothing survives in the 126 MB L2
    launches = int(L.launch_count() - n0)
    path = int(L.jacobi2d_last_path())
    if path != 3:
        raise RuntimeError("headline did not take the marching kernel (jacobi2d_last_path = %d)" % path)
    ms = float(np.mean(times))
    value = units / (ms * 1e-3) / 1e9
    # the dominant kernel's launches per step = passes over memory (the border-ring copy that prepares the scratch
    # grid of the closing two-state pass is a launch, but not a pass)
    per_step = int(L.jacobi2d_last_passes()) or max(1, launches // max(1, args.steps))
    alg = 16.0 * units / per_step
    us = ms * 1e3 / per_step
    ach = alg / (us * 1e-6) / 1e9
    tr = traffic.get("jacobi_2d_weak_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "jacobi2d_march_kernel: 3/5/7 sweeps per pass in registers (csrc/jacobi2d_march.cuh)",
                "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": tr,
                "frac_dram": round(tr / (us * 1e-6) / 1e9 / peak, 4) if tr else None,
                "traffic_source": traffic.get("jacobi_2d_weak_source"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "launches_per_step": per_step, "avg_launch_us": round(us, 2),
                "note": "algorithmic = 16 B per cell update (BASELINE.md section 2); a pass fuses up to 7 sweeps per DRAM "
                        "round trip, so `frac` (algorithmic) exceeds 1 while `frac_dram` (ncu DRAM bytes / time / peak) is what "
                        "the memory system actually sustains"}

    # parity: one step from the initial state, bands at the top border and mid-grid
    init(); step(); L.sync()

    def fetch(lo, hi):
        a = np.empty((hi - lo, WEAK_COLS)); b = np.empty((hi - lo, WEAK_COLS))
        L.d2h(a.ctypes.data, A.ptr + lo * WEAK_COLS * 8, a.nbytes)
        L.d2h(b.ctypes.data, B.ptr + lo * WEAK_COLS * 8, b.nbytes)
        L.sync()
        return a, b
    parity = {"method": "rows of the device result after one step vs the CPU oracle on the same rows + dependency cone",
              "bands": [band_parity(fetch, WEAK_ROWS, c) for c in (8, WEAK_ROWS // 2, WEAK_ROWS - 8)]}
    parity["bit_exact"] = all(b["bit_exact"] for b in parity["bands"])

    # e2e: pinned host arrays through the public host-buffer API
    e2e = None
    try:
        hA, pA = pinned_array(L, (WEAK_ROWS, WEAK_COLS)); hB, pB = pinned_array(L, (WEAK_ROWS, WEAK_COLS))
        L.d2h(hA.ctypes.data, A.ptr, hA.nbytes); L.d2h(hB.ctypes.data, B.ptr, hB.nbytes); L.sync()
        nb.jacobi_2d(WEAK_TSTEPS, hA, hB)
        # the host-buffer call (pipelined over row chunks) against the device-array call on the same state, bit for bit
        nb.jacobi_2d(WEAK_TSTEPS, A, B); L.sync()
        same = True
        for r0 in (0, WEAK_ROWS // 3 - 5, WEAK_ROWS // 2 + 120, WEAK_ROWS - 16):
            a, b = fetch(r0, r0 + 16)
            same = same and np.array_equal(a, hA[r0:r0 + 16]) and np.array_equal(b, hB[r0:r0 + 16])
        tt = []
        for _ in range(3):
            t0 = time.perf_counter(); nb.jacobi_2d(WEAK_TSTEPS, hA, hB); tt.append(time.perf_counter() - t0)
        e = float(np.mean(tt))
        ring = 8 * (2 * WEAK_COLS + 2 * WEAK_ROWS)
        e2e = {"value": round(units / e / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": hA.nbytes + ring,
               "d2h_bytes_per_step": 2 * hA.nbytes, "ms_per_step": round(e * 1e3, 2), "steps": len(tt),
               "matches_device_path": bool(same),
               "api": "npbench_b200.jacobi_2d(TSTEPS, A, B) on pinned host ndarrays -> npb_jacobi2d_f64_host: a pipeline "
                      "over 256-row chunks (H2D of A and of B's border ring -- B's interior is dead on entry --, the 8 "
                      "marching passes skewed one chunk apart, D2H of A and B; three streams)"}
        L.host_free(pA); L.host_free(pB)
        del hA, hB
    except Exception as ex:
        e2e = {"value": None, "unit": UNIT, "note": "failed: %s" % repr(ex)[:200]}
    del A, B
    L.pool_trim()

    try:
        cpu = cpu_baseline_numpy() if find_checkout() else cpu_port()
    except Exception as ex:
        cpu = {"value": None, "unit": UNIT, "note": repr(ex)[:200]}
    try:
        port = cpu_port()
    except Exception as ex:
        port = {"note": repr(ex)[:200]}
    records = run_records(nb, peak, peak_src, max(20, args.steps), not args.no_harness) if not args.no_records else None
    suite = run_suite(nb, peak) if not args.no_suite else None
    line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": warm, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": DATA,
            "config": {"workload": workload_label(1),
                       "l2": "inputs (13.4 GB) far larger than L2; no flush needed",
                       "timing": "CUDA events on the launch stream around each step, mean of K steps",
                       "api": "npbench_b200.jacobi_2d(TSTEPS, A, B) on device arrays -> npb_jacobi2d_f64"},
            "roofline": roofline, "cpu_baseline": cpu, "cpu_port": port, "e2e": e2e, "parity": parity,
            "gpu_launches": launches, "clocks": clk.summary(), "records": records, "suite": suite}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ N > 1
def multi_gpu_suite(args, world, rank, eng, D, torch, dist, peak):
    """Other sharded kernels at N GPUs (max over ranks, CUDA events): fdtd_2d and heat_3d weak-scaled with halo
    exchange (marching passes between exchanges), each beside the no-exchange rate of ONE slab measured in the
    same run; hdiff / vadv `paper` column-sharded (no exchange)."""
    L = eng.lib
    rows = []

    def timed(fn, reps=3, solo=False):
        ts = []
        for _ in range(reps + 1):
            if not solo:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(ts[1:]))], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def add(kernel, workload, units, bpu, ms, solo_units=None, solo_ms=None):
        gc = units / (ms * 1e-3) / 1e9
        row = {"kernel": kernel, "workload": workload, "ms": round(ms, 4), "value": round(gc, 2),
               "frac_of_peak_per_gpu": round(gc * bpu / world / peak, 3)}
        if solo_ms:
            sv = solo_units / (solo_ms * 1e-3) / 1e9
            row["single_gpu_same_workload"] = round(sv, 2)
            row["weak_scaling_efficiency"] = round(gc / (world * sv), 4)
        rows.append(row)

    try:   # fdtd_2d: (N*8192) x 65536, TMAX = 20, ghost depth = steps per marching pass
        nx, ny, tmax, H = world * 8192, 65536, 20, D.FDTD_GHOST
        slab = D.Slab(nx, world, rank, H)
        f = [eng.empty(slab.nloc, ny) for _ in range(3)]
        L.init_fdtd2d_f64(tmax, nx, ny, slab.row0, slab.nloc, f[0].data_ptr(), f[1].data_ptr(), f[2].data_ptr(), 0)
        fict = [float(t) for t in range(tmax)]
        ms = timed(lambda: D.fdtd_2d_sharded(eng, slab, tmax, f[0], f[1], f[2], fict))
        solo = D.Slab(slab.nloc, 1, 0, H)
        sms = timed(lambda: D.fdtd_2d_sharded(eng, solo, tmax, f[0], f[1], f[2], fict), solo=True)
        add("fdtd_2d", "%dx%d TMAX=%d, row slabs, halo every %d steps" % (nx, ny, tmax, H), tmax * nx * ny, 48.0, ms,
            tmax * slab.nloc * ny, sms)
        del f
    except Exception as e:
        rows.append({"kernel": "fdtd_2d", "error": repr(e)[:200]})
    torch.cuda.empty_cache()
    try:   # heat_3d: (N*512) x 1024 x 1024, TSTEPS = 7 (12 sweeps), ghost depth = sweeps per marching pass
        n1, ts, H = 1024, 7, D.HEAT_GHOST
        n0 = world * 512
        slab = D.Slab(n0, world, rank, H)
        A, B = eng.empty(slab.nloc, n1, n1), eng.empty(slab.nloc, n1, n1)
        A.uniform_(); B.copy_(A)
        ms = timed(lambda: D.heat_3d_sharded(eng, slab, ts, A, B))
        solo = D.Slab(slab.nloc, 1, 0, H)
        sms = timed(lambda: D.heat_3d_sharded(eng, solo, ts, A, B), solo=True)
        add("heat_3d", "%dx%dx%d TSTEPS=%d, i-plane slabs, halo every %d sweeps" % (n0, n1, n1, ts, H),
            2 * (ts - 1) * (n0 - 2) * (n1 - 2) ** 2, 16.0, ms, 2 * (ts - 1) * (slab.nloc - 2) * (n1 - 2) ** 2, sms)
        del A, B
    except Exception as e:
        rows.append({"kernel": "heat_3d", "error": repr(e)[:200]})
    torch.cuda.empty_cache()
    try:   # hdiff / vadv `paper`, split along I (fixed overlaps, no run-time exchange)
        I, J, K = 256, 256, 160
        lo, hi = D.hdiff_shard(I, world, rank)
        inf, out, cf = eng.empty(hi - lo + 4, J + 4, K).uniform_(), eng.empty(hi - lo, J, K), eng.empty(hi - lo, J, K).uniform_()
        ms = timed(lambda: eng.hdiff(inf, out, cf), reps=10)
        add("hdiff", "paper 256x256x160 split along I over %d GPUs, no exchange" % world, I * J * K,
            8.0 * ((I + 4) * (J + 4) + 2 * I * J) / (I * J), ms)
        lo, hi = D.vadv_shard(I, world, rank)
        t = [eng.empty(hi - lo, J, K).uniform_() for _ in range(2)] + [eng.empty(hi - lo + 1, J, K).uniform_()] + \
            [eng.empty(hi - lo, J, K).uniform_() for _ in range(2)]
        ms = timed(lambda: eng.vadv(*t, 0.15), reps=10)
        add("vadv", "paper 256x256x160 split along I over %d GPUs, no exchange" % world, I * J * K, 8.0 * (6 * I + 1) / I, ms)
    except Exception as e:
        rows.append({"kernel": "hdiff/vadv", "error": repr(e)[:200]})
    return rows


def multi_gpu(args, world, rank, local_rank):
    import torch
    import torch.distributed as dist
    from npbench_b200 import distributed as D
    # NCCL prints its version banner on stdout at communicator creation; stdout must carry exactly
    # one JSON line, so everything but the final print goes to stderr (fd-level redirect).
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = D.B200Engine(local_rank)
    L = eng.lib
    peak, peak_src = measured_peak()
    traffic = traffic_table()
    n_rows = world * WEAK_ROWS
    slab = D.Slab(n_rows, world, rank, D.JACOBI_MAX_BLOCK)
    A, B = eng.empty(slab.nloc, WEAK_COLS), eng.empty(slab.nloc, WEAK_COLS)

    def init():
        L.init_jacobi2d_f64(WEAK_COLS, slab.row0, slab.nloc, WEAK_COLS, A.data_ptr(), B.data_ptr())

    def step():
        D.jacobi_2d_sharded(eng, slab, WEAK_TSTEPS, A, B)

    init()
    warm = max(3, args.warmup)
    for _ in range(warm):
        step()
    torch.cuda.synchronize(); dist.barrier()
    n0 = L.launch_count()
    evs = []
    with ClockSampler(local_rank) as clk:
        for _ in range(args.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier(); torch.cuda.synchronize()
            e0.record(); step(); e1.record()
            torch.cuda.synchronize()
            evs.append(e0.elapsed_time(e1))
    launches = int(L.launch_count() - n0)
    t = torch.tensor(evs, dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                   # max over ranks, per step
    ms = float(t.mean().item())
    units = weak_units(world)
    value = units / (ms * 1e-3) / 1e9

    # parity: one step from the initial state; every rank checks a band straddling its upper slab seam
    # (rank 0: the global top border), i.e. rows produced from exchanged halos
    init(); step(); torch.cuda.synchronize()

    def fetch(lo, hi):
        a = A[lo - slab.row0:hi - slab.row0].cpu().numpy(); b = B[lo - slab.row0:hi - slab.row0].cpu().numpy()
        return a, b
    centre = max(slab.lo + PARITY_BAND // 2, slab.lo) if rank == 0 else slab.lo + PARITY_BAND // 2
    try:
        mine = band_parity(fetch, n_rows, centre)
    except Exception as ex:
        mine = {"rows": None, "bit_exact": False, "error": repr(ex)[:160]}
    flag = torch.tensor([1.0 if mine["bit_exact"] else 0.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    parity = {"method": "per rank: 16 owned rows next to its slab seam after one step vs the CPU oracle on those rows + "
                        "dependency cone (rows computed from exchanged halos)",
              "rank0_band": mine, "bit_exact": bool(flag.item() == 1.0), "ranks_checked": world}

    # same slab without any exchange (single-GPU rate of this workload, measured here)
    solo = D.Slab(slab.nloc, 1, 0, D.JACOBI_MAX_BLOCK)
    solo_t = []
    for _ in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); D.jacobi_2d_sharded(eng, solo, WEAK_TSTEPS, A, B); e1.record()
        torch.cuda.synchronize()
        solo_t.append(e0.elapsed_time(e1))
    solo_ms = torch.tensor([float(np.median(solo_t))], dtype=torch.float64, device="cuda")
    dist.all_reduce(solo_ms, op=dist.ReduceOp.MAX)
    solo_val = WEAK_SWEEPS * (slab.nloc - 2) * (WEAK_COLS - 2) / (solo_ms.item() * 1e-3) / 1e9

    # e2e: pinned host slabs -> H2D -> sharded kernel -> D2H of the owned rows of A and B
    e2e = None
    try:
        import psutil
        need = 2 * slab.nloc * WEAK_COLS * 8
        if need * world < 0.25 * psutil.virtual_memory().available:
            hA = torch.empty((slab.nloc, WEAK_COLS), dtype=torch.float64, pin_memory=True)
            hB = torch.empty((slab.nloc, WEAK_COLS), dtype=torch.float64, pin_memory=True)
            hA.copy_(A); hB.copy_(B)
            ts = []
            for i in range(3):
                dist.barrier(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                A.copy_(hA, non_blocking=True)
                # B's interior is dead on entry (the first sweep overwrites it): only its border ring goes up
                B[0].copy_(hB[0], non_blocking=True); B[-1].copy_(hB[-1], non_blocking=True)
                B[:, 0].copy_(hB[:, 0], non_blocking=True); B[:, -1].copy_(hB[:, -1], non_blocking=True)
                step()
                slab.owned(hA).copy_(slab.owned(A), non_blocking=True)
                slab.owned(hB).copy_(slab.owned(B), non_blocking=True)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            tt = torch.tensor([min(ts[1:])], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            own = (slab.hi - slab.lo) * WEAK_COLS * 8
            up = (slab.nloc * WEAK_COLS + 2 * WEAK_COLS + 2 * slab.nloc) * 8
            e2e = {"value": round(units / tt.item() / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": up * world,
                   "d2h_bytes_per_step": 2 * own * world, "ms_per_step": round(tt.item() * 1e3, 3),
                   "api": "npbench_b200.distributed.jacobi_2d_sharded on pinned host slabs (H2D of A and of B's border ring "
                          "-- B's interior is dead on entry -- + kernels + NCCL halos + D2H of the owned rows of A and B)"}
            del hA, hB
        else:
            e2e = {"value": None, "unit": UNIT, "note": "skipped: pinned host slabs would not fit comfortably"}
    except Exception as e:
        e2e = {"value": None, "unit": UNIT, "note": "failed: %s" % str(e)[:160]}

    del A, B
    torch.cuda.empty_cache()
    suite = multi_gpu_suite(args, world, rank, eng, D, torch, dist, peak) if not args.no_suite else None

    # dominant kernel: the multi-sweep jacobi pass (up to 7 sweeps fused): 16 B x cells x sweeps per launch
    n_pass = max(1, len(D.jacobi_plan_dual(WEAK_SWEEPS) if eng.jacobi_dual_ok(slab.nloc, WEAK_COLS) else D.jacobi_plan(WEAK_SWEEPS)))
    achieved = value * 16.0 / world
    tr = traffic.get("jacobi_2d_weak_bytes_per_launch")
    pass_us = ms * 1e3 / n_pass
    clocks = clk.summary()
    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": round(ms, 4), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": DATA,
                "config": {"workload": workload_label(world),
                           "sharding": "%d row slabs, ghost depth 7, halo exchange every marching pass (ncclSend/Recv via "
                                       "torch.distributed), overlapped with interior rows" % world,
                           "l2": "inputs (13.4 GB per GPU) far larger than L2; no flush needed",
                           "timing": "CUDA events per step, barrier + synchronize before each, max over ranks"},
                "roofline": {"bound": "hbm", "kernel": "jacobi2d_march_kernel: 3/5/7 sweeps per pass in registers over the slab's row ranges "
                                       "(csrc/jacobi2d_march.cuh, reached through npb_jacobi2d_block_f64)",
                             "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(achieved / peak, 4), "traffic": tr,
                             "frac_dram": round(tr / (pass_us * 1e-6) / 1e9 / peak, 4) if tr else None,
                             "traffic_source": traffic.get("jacobi_2d_weak_source"), "peak_source": peak_src,
                             "note": "per GPU; algorithmic 16 B per cell update, %d passes per step (%.1f sweeps fused per launch); "
                                     "traffic = ncu DRAM bytes of one pass over one 10240-row slab (same kernel, same slab "
                                     "size at every N)" % (n_pass, WEAK_SWEEPS / n_pass)},
                "single_gpu_same_workload": {"value": round(solo_val, 3), "unit": UNIT,
                                             "note": "one slab, no halo exchange, same run; efficiency = value / (n_gpus x this)"},
                "cpu_baseline": {"value": None, "unit": UNIT, "cores": None, "kind": "reference",
                                 "sample": "reported on the N = 1 line only (rank 0 at N = 1, as the contract asks); the "
                                           "--impl reference arm times NumPy at every N"},
                "e2e": e2e, "parity": parity, "gpu_launches": launches, "clocks": clocks, "suite": suite}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-suite", action="store_true")
    ap.add_argument("--no-records", action="store_true")
    ap.add_argument("--no-harness", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.steps is None:
        args.steps = 20 if world == 1 else 5
    if args.impl == "reference":
        reference_arm(args, max(world, args.gpus), rank)
        return
    if world == 1:
        if args.gpus != 1:
            sys.stderr.write("bench.py: --gpus %d needs torchrun (one rank per GPU); running the 1-GPU workload\n" % args.gpus)
        single_gpu(args)
    else:
        multi_gpu(args, world, rank, local_rank)


if __name__ == "__main__":
    main()
