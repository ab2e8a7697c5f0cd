"""Per-group timeline of the streaming vadv solver (profiling aid).
slots: 0 group start, 1 forward end, 2 backward end (globaltimer ns); 3/4 cycles spent waiting for
stages in the forward/backward sweep; 5 = smid<<8 | warp.   usage: [mode] (default 3)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npbench_b200 as nb
nb.init(0)
L = nb.lib()
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 3
I, J, K = 256, 256, 160
rng = np.random.default_rng(0)
a = [nb.DeviceArray.from_host(rng.random(s)) for s in ((I, J, K), (I, J, K), (I + 1, J, K), (I, J, K), (I, J, K))]
ng = (I * J + 31) // 32
tr = nb.DeviceArray((ng * 8,))
L.vadv_set_mode(mode)
for _ in range(3): nb.vadv(*a, 0.15)
L.memset(tr.ptr, 0, ng * 64)
L.vadv_set_trace(tr.ptr)
L.l2_flush(); nb.vadv(*a, 0.15); L.sync()
L.vadv_set_trace(None)
assert L.vadv_last_path() == 3
t = tr.to_host().view(np.uint64).reshape(ng, 8).astype(np.int64)
t0 = t[:, 0].min()
st, fe, be = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3
clk = 1.965e3   # cycles per us
print("mode %d groups %d span %.1f us" % (mode, ng, be.max()))
print("per group: forward %.2f us (waiting for stages %.2f, of which first stage %.2f), backward %.2f us" % (
    (fe - st).mean(), (t[:, 3] / clk).mean(), (t[:, 6] / clk).mean(), (be - fe).mean()))
n_lat = (t[:, 7] >> 32); asm_ns = t[:, 7] & 0xffffffff
print("fills the helper had to wait for: %.1f per group, mean issue->full latency %.2f us; helper assembly %.2f us per group" % (
    n_lat.mean(), (t[:, 4].sum() / max(n_lat.sum(), 1)) / 1e3, asm_ns.mean() / 1e3))
print("forward quantiles us:", np.round(np.quantile(fe - st, [0, .1, .5, .9, 1]), 2))
print("backward quantiles us:", np.round(np.quantile(be - fe, [0, .1, .5, .9, 1]), 2))
sm = t[:, 5] >> 8
w = t[:, 5] & 255
for s in (int(sm.min()), int(np.median(sm))):
    for ww in range(int(w.max()) + 1):
        idx = np.where((sm == s) & (w == ww))[0]
        idx = idx[np.argsort(st[idx])]
        print("SM %d warp %d:" % (s, ww), "  ".join("%.1f-%.1f-%.1f" % (st[i], fe[i], be[i]) for i in idx))
ends = np.array([be[sm == s].max() for s in np.unique(sm)])
print("per-SM finish time us: min %.1f median %.1f max %.1f" % (ends.min(), np.median(ends), ends.max()))
