"""Quick device-timed throughput table for all kernels/presets (development aid, GPU only)."""
import ctypes
import json
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npbench_b200 as nb

PRESETS = {
    "jacobi_2d": {"S": (50, 150), "M": (80, 350), "L": (200, 700), "paper": (1000, 2800), "big": (21, 16384), "n4k": (41, 4096), "n6k": (41, 6144), "n8k": (21, 8192)},
    "heat_3d": {"S": (25, 25), "M": (50, 40), "L": (100, 70), "paper": (500, 120), "big": (11, 640),
                "n160": (41, 160), "n200": (41, 200), "n256": (21, 256), "n384": (11, 384), "n512": (11, 512), "n1024": (5, 1024)},
    "fdtd_2d": {"S": (20, 200, 220), "M": (60, 400, 450), "L": (150, 800, 900), "paper": (500, 1000, 1200),
                "big": (10, 8192, 16384), "n2k": (20, 2048, 2048), "n4k": (20, 4096, 4096), "n3k": (20, 3072, 3072)},
    "hdiff": {"S": (64, 64, 60), "M": (128, 128, 160), "L": (384, 384, 160), "paper": (256, 256, 160)},
    "vadv": {"S": (60, 60, 40), "M": (112, 112, 80), "L": (180, 180, 160), "paper": (256, 256, 160)},
    "jacobi_1d": {"S": (800, 3200), "M": (3000, 12000), "L": (8500, 34000), "paper": (4000, 32000)},
    "seidel_2d": {"S": (8, 50), "M": (15, 100), "L": (40, 200), "paper": (100, 400)},
    "adi": {"S": (5, 100), "M": (20, 200), "L": (50, 500), "paper": (100, 200)},
    "cavity_flow": {"S": (61, 25, 5), "M": (121, 50, 10), "L": (201, 100, 20), "paper": (101, 700, 50)},
    "channel_flow": {"S": (61, 5), "M": (121, 10), "L": (201, 20), "paper": (101, 50)},
}


def timed(fn, reps, flush=True):
    L = nb.lib()
    ms = ctypes.c_float()
    ts = []
    for _ in range(3):
        fn()
    L.sync()
    for _ in range(reps):
        if flush:
            L.l2_flush()
        L.timer_start()
        fn()
        L.timer_stop(ctypes.byref(ms))
        ts.append(ms.value)
    return float(np.median(ts)), float(np.min(ts))


def main():
    nb.init(0)
    L = nb.lib()
    if os.environ.get('NPB_VADV_MODE'):
        L.vadv_set_mode(int(os.environ['NPB_VADV_MODE']))
    if os.environ.get('NPB_J2_MODE'):
        L.jacobi2d_set_mode(int(os.environ['NPB_J2_MODE']))
    if os.environ.get('NPB_FDTD_MODE'):
        L.fdtd2d_set_mode(int(os.environ['NPB_FDTD_MODE']))
    if os.environ.get('NPB_HEAT_MODE'):
        L.heat3d_set_mode(int(os.environ['NPB_HEAT_MODE']))
    rng = np.random.default_rng(0)
    rows = []
    only = sys.argv[1:]
    for bench, presets in PRESETS.items():
        if only and not any(o.split(":")[0] == bench for o in only):
            continue
        for pname, p in presets.items():
            if only and not any(o == bench or o == bench + ":" + pname for o in only):
                continue
            if bench == "jacobi_2d":
                ts, n = p
                A = nb.DeviceArray((n, n)); B = nb.DeviceArray((n, n))
                L.init_jacobi2d_f64(n, 0, n, n, A.ptr, B.ptr)
                fn = lambda: nb.jacobi_2d(ts, A, B)
                units = 2 * (ts - 1) * (n - 2) ** 2; bpu = 16
            elif bench == "heat_3d":
                ts, n = p
                A = nb.DeviceArray((n, n, n)); B = nb.DeviceArray((n, n, n))
                L.init_heat3d_f64(n, 0, n, A.ptr, B.ptr)
                fn = lambda: nb.heat_3d(ts, A, B)
                units = 2 * (ts - 1) * (n - 2) ** 3; bpu = 16
            elif bench == "fdtd_2d":
                tm, nx, ny = p
                a = [nb.DeviceArray((nx, ny)) for _ in range(3)] + [nb.DeviceArray((tm,))]
                L.init_fdtd2d_f64(tm, nx, ny, 0, nx, *(x.ptr for x in a))
                fn = lambda: nb.fdtd_2d(tm, *a)
                units = tm * nx * ny; bpu = 48
            elif bench == "jacobi_1d":
                ts, n = p
                A = nb.DeviceArray.from_host((np.arange(n) + 2.0) / n); B = nb.DeviceArray.from_host((np.arange(n) + 3.0) / n)
                fn = lambda: nb.jacobi_1d(ts, A, B)
                units = 2 * (ts - 1) * (n - 2); bpu = 16
            elif bench == "seidel_2d":
                ts, n = p
                A = nb.DeviceArray.from_host(rng.random((n, n)))
                fn = lambda: nb.seidel_2d(ts, n, A)
                units = (ts - 1) * (n - 2) ** 2; bpu = 16
            elif bench == "adi":
                ts, n = p
                A0 = nb.DeviceArray.from_host(rng.random((n, n))); A = nb.DeviceArray((n, n))
                fn = lambda: (L.d2d(A.ptr, A0.ptr, n * n * 8), nb.adi(ts, n, A))    # the reference scheme diverges: restart every call
                units = 2 * ts * (n - 2) ** 2; bpu = 16
            elif bench == "cavity_flow":
                n, nt, nit = p
                z = np.zeros((n, n)); a0 = [nb.DeviceArray.from_host(z) for _ in range(3)]; a = [nb.DeviceArray((n, n)) for _ in range(3)]
                dx = 2 / (n - 1); dt = .1 / ((n - 1) * (n - 1))

                def fn(a=a, a0=a0, n=n, nt=nt, nit=nit, dx=dx, dt=dt):
                    for x, x0 in zip(a, a0):
                        L.d2d(x.ptr, x0.ptr, n * n * 8)
                    nb.cavity_flow(n, n, nt, nit, a[0], a[1], dt, dx, dx, a[2], 1.0, 0.1)
                units = nt * (nit + 2) * (n - 2) ** 2; bpu = 16
            elif bench == "channel_flow":
                n, nit = p
                a0 = [nb.DeviceArray.from_host(x) for x in (np.zeros((n, n)), np.zeros((n, n)), np.ones((n, n)))]
                a = [nb.DeviceArray((n, n)) for _ in range(3)]
                dx = 2 / (n - 1); dt = .1 / ((n - 1) * (n - 1))
                box = {}

                def fn(a=a, a0=a0, n=n, nit=nit, dx=dx, dt=dt, box=box):
                    for x, x0 in zip(a, a0):
                        L.d2d(x.ptr, x0.ptr, n * n * 8)
                    box["steps"] = nb.channel_flow(nit, a[0], a[1], dt, dx, dx, a[2], 1.0, 0.1, 1.0)
                fn()
                units = box["steps"] * (nit + 2) * n * (n - 2); bpu = 16
            elif bench == "hdiff":
                I, J, K = p
                a = [nb.DeviceArray.from_host(rng.random(s)) for s in ((I + 4, J + 4, K), (I, J, K), (I, J, K))]
                fn = lambda: nb.hdiff(*a)
                units = I * J * K; bpu = 8.0 * ((I + 4) * (J + 4) + 2 * I * J) / (I * J)
            else:
                I, J, K = p
                a = [nb.DeviceArray.from_host(rng.random(s)) for s in
                     ((I, J, K), (I, J, K), (I + 1, J, K), (I, J, K), (I, J, K))]
                fn = lambda: nb.vadv(*a, 0.15)
                units = I * J * K; bpu = 8.0 * (6 * I + 1) / I
            reps = 3 if pname in ("paper", "big") and bench in ("jacobi_2d", "heat_3d", "fdtd_2d") else 10
            n0 = L.launch_count()
            med, best = timed(fn, reps)
            launches = (L.launch_count() - n0) // (reps + 3)
            gc = units / (med * 1e-3) / 1e9
            row = dict(bench=bench, preset=pname, params=p, ms=round(med, 4), best_ms=round(best, 4),
                       gcell_s=round(gc, 2), eff_gbs=round(gc * bpu, 1), frac_6553=round(gc * bpu / 6553.6, 3),
                       launches=int(launches))
            rows.append(row)
            print(json.dumps(row), flush=True)
            del fn
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/quick_perf.json", "w"), indent=1)


if __name__ == "__main__":
    main()
