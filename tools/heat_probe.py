"""heat_3d resident kernels: per-sweep time (slope between two TSTEPS) and fixed cost (intercept), per mode.

    python tools/heat_probe.py [N ...]          # development aid, GPU only
Modes: 6 register-tile kernel, 2 shared-memory resident kernel (round 1); +256 no fences, +512 no polls,
+1024 no sends (timing experiments of the register-tile kernel: wrong results by construction).
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npbench_b200 as nb


def timed(L, fn, reps=15, flush=True):
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
    return float(np.median(ts)) * 1e3


def main():
    nb.init(0)
    L = nb.lib()
    sizes = [int(a) for a in sys.argv[1:]] or [25, 40, 70]
    modes = [int(m) for m in os.environ.get("MODES", "6,2,262,518,1542,1798").split(",")]
    t1, t2 = 26, 126
    for n in sizes:
        A = nb.DeviceArray((n, n, n)); B = nb.DeviceArray((n, n, n))
        L.init_heat3d_f64(n, 0, n, A.ptr, B.ptr)
        for m in modes:
            L.heat3d_set_mode(m)
            try:
                a = timed(L, lambda: nb.heat_3d(t1, A, B))
                b = timed(L, lambda: nb.heat_3d(t2, A, B))
                c = timed(L, lambda: nb.heat_3d(t2, A, B), flush=False)
            except Exception as e:
                print("N=%d mode=%d: %s" % (n, m, e))
                continue
            finally:
                L.heat3d_set_mode(0)
            per = (b - a) / (2 * (t2 - t1))
            print("N=%3d mode=%4d path=%d: %6.3f us/sweep, fixed %6.1f us (TSTEPS=%d: %.1f us; %d: %.1f us; no flush %.1f us)" % (
                n, m, L.heat3d_last_path(), per, a - per * 2 * (t1 - 1), t1, a, t2, b, c))


if __name__ == "__main__":
    main()
