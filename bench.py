#!/usr/bin/env python
"""bench.py -- throughput of the B200 stencil hot path, one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): Gcell-updates/s and fraction of the HBM roofline.

N = 1   headline workload = BASELINE.json configs[1]: heat_3d preset L (TSTEPS=100, N=70,
        fp64).  One *step* = one call kernel(TSTEPS, A, B) = 198 sweeps.  Arrays are resident
        in HBM; L2 is flushed (untimed) before every timed step; each step is timed with
        CUDA events on the launch stream.  The line also carries
          roofline      dominant kernel vs the measured HBM peak (MEASURED_PEAKS.json),
          e2e           the same call through the public host-buffer API (pinned host
                        arrays, H2D + kernels + D2H inside the timed region),
          cpu_baseline  the CPU oracle port on this box's host cores (bounded sample),
          suite         every kernel x NPBench preset + the scaled single-GPU grids,
          clocks        SM clock / throttle reasons sampled via NVML during the timed region.
N > 1   headline workload = BASELINE.json configs[4]: jacobi_2d on a weak-scaled grid
        ((N*10240) x 81920 fp64, TSTEPS=21), row slabs, halo exchange over NCCL overlapped
        with interior compute (npbench_b200/distributed.py).  `single_gpu_same_workload` in
        the line is the no-exchange rate of one slab, measured in the same run, so scaling
        efficiency can be read without mixing workloads.
--impl reference   the CPU arm: the reference algorithm (oracle port of the NumPy functions,
        all host threads) on the same config/metric; rank 0 only.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gcell-updates/s (fp64 stencil cell updates per second)"
UNIT = "Gcell/s"
HEAT_L = dict(TSTEPS=100, N=70)
WEAK_ROWS, WEAK_COLS, WEAK_TSTEPS = 10240, 81920, 21


def measured_peak():
    """HBM roofline denominator: MEASURED_PEAKS.json (driver-written) else the recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


# ------------------------------------------------------------------ CPU arm (oracle port)
def cpu_heat3d_L(budget_s, threads):
    import oracle
    oracle.set_threads(threads)
    A, B = oracle.init_heat_3d(HEAT_L["N"])
    units, _ = units_of("heat_3d", HEAT_L)
    oracle.heat_3d(HEAT_L["TSTEPS"], A, B)          # warm-up
    times, t_end = [], time.perf_counter() + budget_s
    while len(times) < 3 or (time.perf_counter() < t_end and len(times) < 50):
        t0 = time.perf_counter()
        oracle.heat_3d(HEAT_L["TSTEPS"], A, B)
        times.append(time.perf_counter() - t0)
    t = float(np.median(times))
    return {"value": round(units / t / 1e9, 4), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "full workload (heat_3d L, 198 sweeps of 70^3), median of %d runs of the C/OpenMP oracle port "
                      "(oracle/stencil_oracle.c, NumPy evaluation order, -ffp-contract=off); NumPy itself uses 1 core "
                      "for this kernel" % len(times),
            "ms_per_step": round(t * 1e3, 3)}


def reference_arm(args, world, rank):
    """--impl reference: the reference algorithm on host cores, same metric/config."""
    if rank != 0:
        return
    import oracle
    threads = oracle.max_threads()
    oracle.set_threads(threads)
    if world == 1:
        workload = "heat_3d preset L (TSTEPS=100, N=70), fp64"
        A, B = oracle.init_heat_3d(HEAT_L["N"])
        units, _ = units_of("heat_3d", HEAT_L)
        step = lambda: oracle.heat_3d(HEAT_L["TSTEPS"], A, B)
        sample = "full workload per step"
    else:
        # bounded sample of the weak-scaled jacobi_2d grid: a 2048-row band of the global grid,
        # same column count and sweep count (per-cell cost is size independent once out of cache)
        rows = 2048
        workload = "jacobi_2d weak-scaled (%d x %d, TSTEPS=%d), fp64, %d row slabs" % (
            world * WEAK_ROWS, WEAK_COLS, WEAK_TSTEPS, world)
        A, B = oracle.init_jacobi_2d(WEAK_COLS, row0=0, nrows=rows, ncols=WEAK_COLS)
        units = 2 * (WEAK_TSTEPS - 1) * (rows - 2) * (WEAK_COLS - 2)
        step = lambda: oracle.jacobi_2d(WEAK_TSTEPS, A, B)
        sample = "%d-row band of the %d-row grid per step (same columns, same TSTEPS)" % (rows, world * WEAK_ROWS)
    for _ in range(max(1, min(args.warmup, 3))):
        step()
    t0 = time.perf_counter(); step(); est = time.perf_counter() - t0
    steps = max(1, min(args.steps, int(120.0 / max(est, 1e-6))))     # keep the arm within ~2 minutes
    times = []
    for _ in range(steps):
        t0 = time.perf_counter(); step(); times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    val = units / t / 1e9
    line = {"metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": round(t * 1e3, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (NPBench initialize closed forms)",
            "impl": "reference", "config": {"workload": workload},
            "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ N = 1
def single_gpu(args):
    import npbench_b200 as nb
    nb.init(0)
    L = nb.lib()
    peak, peak_src = measured_peak()
    p = HEAT_L
    n = p["N"]
    units, bpu = units_of("heat_3d", p)
    A, B = nb.DeviceArray((n, n, n)), nb.DeviceArray((n, n, n))
    L.init_heat3d_f64(n, 0, n, A.ptr, B.ptr)
    step = lambda: nb.heat_3d(p["TSTEPS"], A, B)
    for _ in range(max(3, args.warmup)):
        step()
    L.sync()
    times = []
    n0 = L.launch_count()
    with ClockSampler(0) as clk:
        for _ in range(args.steps):
            times.append(dev_timer(L, step, flush=True))
    launches = int(L.launch_count() - n0)
    ms = float(np.mean(times))
    value = units / (ms * 1e-3) / 1e9

    # dominant kernel.  heat_3d L takes the on-chip resident kernel: ONE launch per step runs all
    # 198 sweeps, so algorithmic bytes per launch = 16 B x interior cells x sweeps.
    per_step_launches = max(1, launches // max(1, args.steps))
    path = {1: "heat3d_resident_kernel (one cooperative launch = all sweeps, tiles resident in shared memory)",
            2: "heat3d_sweep_kernel (one launch per sweep)", 3: "heat3d_tb_kernel (3 sweeps per launch)"}.get(
        int(L.heat3d_last_path()), "heat3d")
    alg_bytes = 16.0 * units / per_step_launches
    avg_launch_us = ms * 1e3 / per_step_launches
    achieved = alg_bytes / (avg_launch_us * 1e-6) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = json.load(f).get("heat_3d_L_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": path + " (csrc/heat3d.cu)", "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "launches_per_step": per_step_launches, "avg_launch_us": round(avg_launch_us, 3),
                "note": "algorithmic = 16 B per cell update (BASELINE.md section 2); the 70^3 grid (2.7 MB/array) stays in "
                        "shared memory for the whole time loop, so DRAM traffic is ~0.6% of the algorithmic bytes and the "
                        "kernel is bound by the per-sweep L2 halo round trip, not by HBM"}

    # e2e: public host-buffer API on pinned NumPy arrays, H2D + 198 sweeps + D2H per step
    import oracle
    hA, pA = pinned_array(L, (n, n, n)); hB, pB = pinned_array(L, (n, n, n))
    a0, b0 = oracle.init_heat_3d(n)
    hA[...] = a0; hB[...] = b0
    for _ in range(3):
        nb.heat_3d(p["TSTEPS"], hA, hB)
    e2e_t = []
    for _ in range(max(5, min(args.steps, 50))):
        L.l2_flush(); L.sync()
        t0 = time.perf_counter()
        nb.heat_3d(p["TSTEPS"], hA, hB)
        e2e_t.append(time.perf_counter() - t0)
    e2e_s = float(np.mean(e2e_t))
    e2e = {"value": round(units / e2e_s / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": 2 * n ** 3 * 8,
           "d2h_bytes_per_step": 2 * n ** 3 * 8, "ms_per_step": round(e2e_s * 1e3, 4),
           "api": "npbench_b200.heat_3d(TSTEPS, A, B) on pinned host ndarrays -> npb_heat3d_f64_host"}
    L.host_free(pA); L.host_free(pB)

    cpu = cpu_heat3d_L(8.0, oracle.max_threads())
    suite = run_suite(nb, peak) if not args.no_suite else None
    line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (NPBench initialize: heat_3d.py:6-11 closed form, generated on device)",
            "config": {"workload": "heat_3d preset L (TSTEPS=100, N=70), fp64, 198 sweeps per step",
                       "l2": "flushed (2x L2-size memset, untimed) before every timed step",
                       "timing": "CUDA events on the launch stream around each step, mean of K steps"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clk.summary(), "suite": suite}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ N > 1
def multi_gpu_suite(args, world, rank, eng, D, torch, dist, peak):
    """Other sharded kernels at N GPUs (max over ranks, CUDA events): fdtd_2d weak-scaled with halo
    exchange; heat_3d weak-scaled with halo exchange; hdiff / vadv `paper` column-sharded (no exchange)."""
    L = eng.lib
    rows = []

    def timed(fn, reps=3):
        ts = []
        for _ in range(reps + 1):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(ts[1:]))], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def add(kernel, workload, units, bpu, ms):
        gc = units / (ms * 1e-3) / 1e9
        rows.append({"kernel": kernel, "workload": workload, "ms": round(ms, 4), "value": round(gc, 2),
                     "frac_of_peak_per_gpu": round(gc * bpu / world / peak, 3)})

    try:   # fdtd_2d: (N*8192) x 65536, TMAX = 10, ghost depth 4
        nx, ny, tmax, H = world * 8192, 65536, 10, 4
        slab = D.Slab(nx, world, rank, H)
        f = [eng.empty(slab.nloc, ny) for _ in range(3)]
        L.init_fdtd2d_f64(tmax, nx, ny, slab.row0, slab.nloc, f[0].data_ptr(), f[1].data_ptr(), f[2].data_ptr(), 0)
        fict = [float(t) for t in range(tmax)]
        ms = timed(lambda: D.fdtd_2d_sharded(eng, slab, tmax, f[0], f[1], f[2], fict))
        add("fdtd_2d", "%dx%d TMAX=%d, row slabs, halo every %d steps" % (nx, ny, tmax, H), tmax * nx * ny, 48.0, ms)
        del f
    except Exception as e:
        rows.append({"kernel": "fdtd_2d", "error": str(e)[:160]})
    torch.cuda.empty_cache()
    try:   # heat_3d: (N*512) x 1024 x 1024, TSTEPS = 6, ghost depth 4
        n1, ts, H = 1024, 6, 4
        n0 = world * 512
        slab = D.Slab(n0, world, rank, H)
        A, B = eng.empty(slab.nloc, n1, n1), eng.empty(slab.nloc, n1, n1)
        A.uniform_(); B.copy_(A)
        ms = timed(lambda: D.heat_3d_sharded(eng, slab, ts, A, B))
        add("heat_3d", "%dx%dx%d TSTEPS=%d, i-plane slabs, halo every %d sweeps" % (n0, n1, n1, ts, H),
            2 * (ts - 1) * (n0 - 2) * (n1 - 2) ** 2, 16.0, ms)
        del A, B
    except Exception as e:
        rows.append({"kernel": "heat_3d", "error": str(e)[:160]})
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
        rows.append({"kernel": "hdiff/vadv", "error": str(e)[:160]})
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
    n_rows = world * WEAK_ROWS
    slab = D.Slab(n_rows, world, rank, D.JACOBI_MAX_BLOCK)
    A, B = eng.empty(slab.nloc, WEAK_COLS), eng.empty(slab.nloc, WEAK_COLS)

    def init():
        L.init_jacobi2d_f64(WEAK_COLS, slab.row0, slab.nloc, WEAK_COLS, A.data_ptr(), B.data_ptr())

    def step():
        D.jacobi_2d_sharded(eng, slab, WEAK_TSTEPS, A, B)

    init()
    for _ in range(max(3, args.warmup)):
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
    units = 2 * (WEAK_TSTEPS - 1) * (n_rows - 2) * (WEAK_COLS - 2)
    value = units / (ms * 1e-3) / 1e9

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
    solo_val = 2 * (WEAK_TSTEPS - 1) * (slab.nloc - 2) * (WEAK_COLS - 2) / (solo_ms.item() * 1e-3) / 1e9

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
                A.copy_(hA, non_blocking=True); B.copy_(hB, non_blocking=True)
                step()
                slab.owned(hA).copy_(slab.owned(A), non_blocking=True)
                slab.owned(hB).copy_(slab.owned(B), non_blocking=True)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t0)
            tt = torch.tensor([min(ts[1:])], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            own = (slab.hi - slab.lo) * WEAK_COLS * 8
            e2e = {"value": round(units / tt.item() / 1e9, 3), "unit": UNIT, "h2d_bytes_per_step": need * world,
                   "d2h_bytes_per_step": 2 * own * world, "ms_per_step": round(tt.item() * 1e3, 3),
                   "api": "npbench_b200.distributed.jacobi_2d_sharded on pinned host slabs (H2D + kernels + NCCL halos + D2H)"}
            del hA, hB
        else:
            e2e = {"value": None, "unit": UNIT, "note": "skipped: pinned host slabs would not fit comfortably"}
    except Exception as e:
        e2e = {"value": None, "unit": UNIT, "note": "failed: %s" % str(e)[:160]}

    del A, B
    torch.cuda.empty_cache()
    suite = multi_gpu_suite(args, world, rank, eng, D, torch, dist, peak) if not args.no_suite else None

    # dominant kernel: the multi-sweep jacobi pass (up to 7 sweeps fused): 16 B x cells x sweeps per launch
    per_launch_sweeps = 2 * (WEAK_TSTEPS - 1) / max(1, len(D.jacobi_plan(2 * (WEAK_TSTEPS - 1))))
    achieved = value * 16.0 / world
    clocks = clk.summary()
    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": round(ms, 4), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic (NPBench initialize: jacobi_2d.py:6-10 closed form, generated on device per slab)",
                "config": {"workload": "jacobi_2d weak-scaled (%d x %d, TSTEPS=%d), fp64, %d row slabs, ghost depth 7, "
                                       "halo exchange every blocked pass (ncclSend/Recv via torch.distributed), "
                                       "overlapped with interior tiles" % (n_rows, WEAK_COLS, WEAK_TSTEPS, world),
                           "l2": "inputs (13.4 GB per GPU) far larger than L2; no flush needed",
                           "timing": "CUDA events per step, barrier + synchronize before each, max over ranks"},
                "roofline": {"bound": "hbm", "kernel": "jacobi2d_march_kernel: 3/5/7 sweeps per pass in registers over the slab's row ranges "
                                       "(csrc/jacobi2d_march.cuh, reached through npb_jacobi2d_block_f64)",
                             "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                             "note": "per GPU, algorithmic 16 B per cell update; ~%.1f sweeps fused per launch so DRAM "
                                     "traffic is far below the algorithmic bytes" % per_launch_sweeps},
                "single_gpu_same_workload": {"value": round(solo_val, 3), "unit": UNIT,
                                             "note": "one slab, no halo exchange, same run; efficiency = value / (n_gpus x this)"},
                "cpu_baseline": None, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "suite": suite}
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
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.steps is None:
        args.steps = 200 if world == 1 else 5
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
