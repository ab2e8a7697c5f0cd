"""Run one kernel x preset a few times (for ncu): python tools/ncu_one.py <bench> <preset|weak> [reps]
weak = the bench.py headline slab (jacobi_2d 10240 x 81920, TSTEPS=21)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import npbench_b200 as nb  # noqa: E402

name, preset = sys.argv[1], sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
nb.init(0)
L = nb.lib()
if preset == "weak":
    p = dict(TSTEPS=bench.WEAK_TSTEPS, NI=bench.WEAK_ROWS, NJ=bench.WEAK_COLS)
else:
    p = [q for b, pr, q in bench.SUITE if b == name and pr == preset][0]
keep, step = bench.make_device_case(nb, name, p, np.random.default_rng(42))
for _ in range(reps):
    L.l2_flush()
    step()
L.sync()
print("done", name, preset)
