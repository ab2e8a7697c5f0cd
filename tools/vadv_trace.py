"""Phase timeline of the vadv pipeline kernel (profiling aid): per-group globaltimer stamps.
slots: 0 A-start, 1 A-end (mover 0), 2 BC-start, 3 BC-end (solver), 4 D-start, 5 D-end (mover 0)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npbench_b200 as nb
nb.init(0)
L = nb.lib()
I, J, K = 256, 256, 160
rng = np.random.default_rng(0)
a = [nb.DeviceArray.from_host(rng.random(s)) for s in ((I, J, K), (I, J, K), (I + 1, J, K), (I, J, K), (I, J, K))]
ng = 4096
tr = nb.DeviceArray((ng * 8,))
for _ in range(3): nb.vadv(*a, 0.15)
L.memset(tr.ptr, 0, ng * 64)
L.vadv_set_trace(tr.ptr)
L.l2_flush(); nb.vadv(*a, 0.15); L.sync()
L.vadv_set_trace(None)
t = tr.to_host().view(np.uint64).reshape(ng, 8)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
d = (t[:, :6].astype(np.int64) - np.int64(t0)) / 1000.0
print("groups", len(t), "span us %.1f" % d[:, 5].max())
print("means us: A %.1f | wait->BC %.1f | BC %.1f | wait->D %.1f | D %.1f" % (
    (d[:,1]-d[:,0]).mean(), (d[:,2]-d[:,1]).mean(), (d[:,3]-d[:,2]).mean(), (d[:,4]-d[:,3]).mean(), (d[:,5]-d[:,4]).mean()))
stride = 3 * 148
for b in (0, 77):
    print("CTA", b)
    for n in range(3):
        for tl in range(3):
            g = b * 3 + tl + n * stride
            print("   n%d t%d: A %.1f-%.1f  BC %.1f-%.1f  D %.1f-%.1f" % (n, tl, d[g,0], d[g,1], d[g,2], d[g,3], d[g,4], d[g,5]))
