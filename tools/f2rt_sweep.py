"""fdtd_2d register-tile resident kernel: configuration sweep (development aid, GPU only).

    python tools/f2rt_sweep.py [S M L | TMAXxNXxNY ...] [--cfgs "rb,cb,nw,T;..."]
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npbench_b200 as nb
import oracle

PRESETS = {"S": (20, 200, 220), "M": (60, 400, 450), "L": (150, 800, 900), "paper": (500, 1000, 1200)}


def timed(fn, reps=20):
    L = nb.lib()
    ms = ctypes.c_float()
    ts = []
    for _ in range(3):
        fn()
    L.sync()
    for _ in range(reps):
        L.l2_flush()
        L.timer_start()
        fn()
        L.timer_stop(ctypes.byref(ms))
        ts.append(ms.value)
    return float(np.median(ts)), float(np.min(ts))


def main():
    args = sys.argv[1:]
    cfgs = []
    if "--cfgs" in args:
        k = args.index("--cfgs")
        cfgs = [c for c in args[k + 1].split(";") if c]
        args = args[:k] + args[k + 2:]
    nb.init(0)
    L = nb.lib()
    oracle.set_threads(oracle.max_threads())
    for pn in args or ["S", "M", "L"]:
        tm, nx, ny = PRESETS[pn] if pn in PRESETS else tuple(int(x) for x in pn.split("x"))
        f0 = oracle.init_fdtd_2d(tm, nx, ny)
        ref = [a.copy() for a in f0]
        oracle.fdtd_2d(tm, *ref)
        units = tm * nx * ny
        for cfg in [None, "per-step"] + cfgs:
            os.environ.pop("NPB_F2R_CFG", None)
            L.fdtd2d_set_mode(1 if cfg == "per-step" else 0)
            if cfg and cfg != "per-step":
                os.environ["NPB_F2R_CFG"] = cfg
            d = [nb.DeviceArray.from_host(a) for a in f0]
            try:
                nb.fdtd_2d(tm, *d)
            except Exception as e:          # noqa: BLE001
                print("%s %-12s ERROR %s" % (pn, cfg, e), flush=True)
                continue
            path = L.fdtd2d_last_path()
            out = (ctypes.c_int * 6)()
            L.fdtd2d_regtile_config(ctypes.cast(out, ctypes.c_void_p))
            ok = all(np.array_equal(d[i].to_host(), ref[i]) for i in range(3))
            d2 = [nb.DeviceArray.from_host(a) for a in f0]
            med, mn = timed(lambda: nb.fdtd_2d(tm, *d2))
            print("%s %-12s path=%d cfg=%s exact=%s  %.4f ms (min %.4f)  %.1f Gcell/s  frac %.3f  %.3f us/step" % (
                pn, cfg, path, list(out) if path == 3 else "-", ok, med, mn, units / med / 1e6,
                units * 48 / med / 1e6 / 6553.6, med * 1e3 / tm), flush=True)
    os.environ.pop("NPB_F2R_CFG", None)
    L.fdtd2d_set_mode(0)


if __name__ == "__main__":
    main()
