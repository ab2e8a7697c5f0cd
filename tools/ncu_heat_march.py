"""One call of heat_3d in mode 5 (three sweeps per pass) at 640^3, for an ncu capture of heat3d_march_kernel."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import npbench_b200 as nb  # noqa: E402

L = nb.lib()
n = 640
A = nb.DeviceArray((n, n, n)); B = nb.DeviceArray((n, n, n))
L.init_heat3d_f64(n, 0, n, A.ptr, B.ptr)
L.heat3d_set_mode(5)
nb.heat_3d(3, A, B)
A.to_host()
