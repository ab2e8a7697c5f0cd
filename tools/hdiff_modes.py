import sys, ctypes, numpy as np
sys.path.insert(0, '.')
import npbench_b200 as nb, oracle
nb.init(0); L = nb.lib()
def timed(fn, reps=30):
    ms = ctypes.c_float(); ts = []
    for _ in range(3): fn()
    L.sync()
    for _ in range(reps):
        L.l2_flush(); L.timer_start(); fn(); L.timer_stop(ctypes.byref(ms)); ts.append(ms.value)
    return np.median(ts) * 1e3
for name, (I, J, K) in {"S": (64, 64, 60), "M": (128, 128, 160), "L": (384, 384, 160), "paper": (256, 256, 160)}.items():
    inf, outf, coeff = oracle.init_hdiff(I, J, K)
    d = [nb.DeviceArray.from_host(a) for a in (inf, outf, coeff)]
    for mode in (0, 1, 2):
        L.hdiff_set_mode(mode)
        t = timed(lambda: nb.hdiff(*d))
        print("hdiff %-5s mode %d path %d: %.1f us  frac %.3f" % (name, mode, L.hdiff_last_path(), t, I * J * K * (24.25 if name != "S" else 25) / t / 1e3 / 6553.6), flush=True)
    L.hdiff_set_mode(0)
