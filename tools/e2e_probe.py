"""Host-buffer (e2e) time of the bench headline workload for several pipeline chunk heights (development aid, GPU only)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import npbench_b200 as nb  # noqa: E402

nb.init(0)
L = nb.lib()
rows, cols, ts = bench.WEAK_ROWS, bench.WEAK_COLS, bench.WEAK_TSTEPS
hA, pA = bench.pinned_array(L, (rows, cols)); hB, pB = bench.pinned_array(L, (rows, cols))
rng = np.random.default_rng(0)
hA[...] = rng.random((1, cols)); hB[...] = hA
units = 2 * (ts - 1) * (rows - 2) * (cols - 2)
for r in [int(x) for x in (sys.argv[1:] or ["64", "128", "256", "512", "0"])]:
    if r:
        os.environ["NPB_J2_PIPE_ROWS"] = str(r)
        os.environ.pop("NPB_J2_PIPE_MIN_CELLS", None)
    else:
        os.environ["NPB_J2_PIPE_MIN_CELLS"] = str(1 << 62)       # pipeline off: plain H2D + passes + D2H
    nb.jacobi_2d(ts, hA, hB)
    tt = []
    for _ in range(3):
        t0 = time.perf_counter(); nb.jacobi_2d(ts, hA, hB); tt.append(time.perf_counter() - t0)
    print("chunk rows %4d: %.1f ms  %.1f Gcell/s" % (r, 1e3 * float(np.mean(tt)), units / float(np.mean(tt)) / 1e9), flush=True)
