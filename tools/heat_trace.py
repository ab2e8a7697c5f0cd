"""Phase cycle counters of heat3d_regtile_kernel's centre CTA (development aid, GPU only).

    python tools/heat_trace.py [N [TSTEPS]]
Prints, for thread 0 (a corner block: two halo sides) and the middle thread, the average cycles per sweep spent in
[faces + halo-independent arithmetic | waiting for halo values | halo-dependent arithmetic + sends | re-arm + publish | fence + barrier],
the %globaltimer timeline of the centre CTA (load | sweep 1 | sweeps 2 - 17 | the rest | store), the spread of the entry and
exit times of all CTAs, and the CUDA-event time of the same call (L2 flushed in front of it, like bench.py does).
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import npbench_b200 as nb


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 70
    ts = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    nb.init(0)
    L = nb.lib()
    A = nb.DeviceArray((n, n, n)); B = nb.DeviceArray((n, n, n))
    L.init_heat3d_f64(n, 0, n, A.ptr, B.ptr)
    tr = nb.DeviceArray((16 + 2 * 148,))
    for mode in [int(m) for m in os.environ.get("MODES", "6,262,518,1798").split(",")]:
        L.heat3d_set_mode(mode)
        L.memset(tr.ptr, 0, 8 * (16 + 2 * 148))
        L.heat3d_set_trace(tr.ptr)
        ms = ctypes.c_float()
        for _ in range(3):
            L.l2_flush()
            L.timer_start()
            nb.heat_3d(ts, A, B)
            L.timer_stop(ctypes.byref(ms))
        print("N=%d mode=%4d CUDA events around the last call: %.2f us" % (n, mode, ms.value * 1e3))
        L.sync()
        L.heat3d_set_trace(None)
        L.heat3d_set_mode(0)
        whole = np.frombuffer(tr.to_host().tobytes(), dtype=np.int64).astype(np.float64)
        raw = whole[:10] / (2 * (ts - 1))
        st = whole[10:16]
        ctas = whole[16:].reshape(-1, 2); ctas = ctas[ctas[:, 0] > 0]
        t0 = ctas[:, 0].min()
        print("N=%d mode=%4d %d CTAs: entry %.2f .. %.2f us after the first, exit %.2f .. %.2f us; centre CTA entry %.2f" % (n, mode, len(ctas), 0.0, (ctas[:, 0].max() - t0) / 1e3, (ctas[:, 1].min() - t0) / 1e3, (ctas[:, 1].max() - t0) / 1e3, (st[0] - t0) / 1e3))
        print("N=%d mode=%4d timeline of the centre CTA (us): load %.2f | sweep 1 %.2f | sweeps 2-17 %.2f | sweeps 18-%d %.2f | store %.2f | total %.2f" % (n, mode, (st[1]-st[0])/1e3, (st[2]-st[1])/1e3, (st[3]-st[2])/1e3, 2*(ts-1), (st[4]-st[3])/1e3, (st[5]-st[4])/1e3, (st[5]-st[0])/1e3))
        for who, v in (("thread 0 (corner block)", raw[:5]), ("middle thread", raw[5:])):
            print("N=%d mode=%4d %-24s: early %6.0f | wait %6.0f | late+send %6.0f | arm+publish %6.0f | fence+barrier %6.0f | sum %6.0f cycles/sweep"
                  % (n, mode, who, v[0], v[1], v[2], v[3], v[4], v.sum()))


if __name__ == "__main__":
    main()
