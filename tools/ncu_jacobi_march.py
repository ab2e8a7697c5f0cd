"""One call of jacobi_2d in mode 3 (marching passes) for an ncu capture of jacobi2d_march_kernel; argv[1] = N (default 16384)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import npbench_b200 as nb  # noqa: E402

nb.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.default_rng(0)
A = nb.DeviceArray.from_host(rng.random((n, n)))
B = nb.DeviceArray.from_host(rng.random((n, n)))
nb.lib().jacobi2d_set_mode(3)
nb.jacobi_2d(9, A, B)
print(nb.lib().jacobi2d_last_path(), float(A.to_host()[5, 5]))
