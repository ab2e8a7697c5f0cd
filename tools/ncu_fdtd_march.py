"""One call of fdtd_2d (TMAX=10, 8192 x 16384: two marching passes of five steps) for an ncu capture of fdtd2d_march_kernel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import npbench_b200 as nb  # noqa: E402

nb.init(0)
nx, ny, tmax = 8192, 16384, 10
rng = np.random.default_rng(0)
f = [nb.DeviceArray.from_host(rng.random((nx, ny))) for _ in range(3)]
fict = nb.DeviceArray.from_host(np.arange(tmax, dtype=np.float64))
nb.fdtd_2d(tmax, f[0], f[1], f[2], fict)
print(nb.lib().fdtd2d_last_path(), float(f[2].to_host()[5, 5]))
