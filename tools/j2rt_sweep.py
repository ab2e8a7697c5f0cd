"""jacobi_2d register-tile resident kernel: configuration sweep (development aid, GPU only).

    python tools/j2rt_sweep.py [S M L ...] [--cfgs "rb,cb,nw,T;rb,cb,nw,T;..."]

For every preset: the default dispatch (model-picked configuration) and each forced configuration (NPB_J2R_CFG) is
checked bit for bit against the oracle once and timed with CUDA events (L2 flushed before every call).
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npbench_b200 as nb
import oracle

PRESETS = {"S": (50, 150), "M": (80, 350), "L": (200, 700), "X": (100, 1000), "T": (30, 60)}


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
    presets = args or ["S", "M", "L"]
    nb.init(0)
    L = nb.lib()
    oracle.set_threads(oracle.max_threads())
    for pn in presets:
        if pn in PRESETS:
            ts, n = PRESETS[pn]
            A0, B0 = oracle.init_jacobi_2d(n)
            nn = (n, n)
        else:                                  # "TSTEPSxNIxNJ", random fields
            ts, ni, nj = (int(x) for x in pn.split("x"))
            rng = np.random.default_rng(3)
            A0, B0 = rng.random((ni, nj)), rng.random((ni, nj))
            nn = (ni, nj)
        rA, rB = A0.copy(), B0.copy()
        oracle.jacobi_2d(ts, rA, rB)
        units = 2 * (ts - 1) * (nn[0] - 2) * (nn[1] - 2)
        for cfg in [None, "blocked"] + cfgs:
            os.environ.pop("NPB_J2R_CFG", None)
            L.jacobi2d_set_mode(0)
            if cfg == "blocked":
                L.jacobi2d_set_mode(1)
            elif cfg:
                os.environ["NPB_J2R_CFG"] = cfg
            dA, dB = nb.DeviceArray.from_host(A0), nb.DeviceArray.from_host(B0)
            try:
                nb.jacobi_2d(ts, dA, dB)
            except Exception as e:          # noqa: BLE001
                print("%s %-14s ERROR %s" % (pn, cfg, e), flush=True)
                continue
            path = L.jacobi2d_last_path()
            out = (ctypes.c_int * 7)()
            L.jacobi2d_regtile_config(ctypes.cast(out, ctypes.c_void_p))
            ok = np.array_equal(dA.to_host(), rA) and np.array_equal(dB.to_host(), rB)
            dA2, dB2 = nb.DeviceArray.from_host(A0), nb.DeviceArray.from_host(B0)
            med, mn = timed(lambda: nb.jacobi_2d(ts, dA2, dB2))
            print("%s %-14s path=%d cfg=%s exact=%s  %.4f ms (min %.4f)  %.1f Gcell/s  frac %.3f  %.3f us/sweep" % (
                pn, cfg, path, list(out) if path == 1 else "-", ok, med, mn, units / med / 1e6,
                units * 16 / med / 1e6 / 6553.6, med * 1e3 / (2 * (ts - 1))), flush=True)
    os.environ.pop("NPB_J2R_CFG", None)
    L.jacobi2d_set_mode(0)


if __name__ == "__main__":
    main()
