# B200-native fdtd_2d: same signature as fdtd_2d_numpy.py:4 (bench_info/fdtd_2d.json input_args).
from npbench_b200 import kernels as _k


def kernel(TMAX, ex, ey, hz, _fict_):
    _k.fdtd_2d(TMAX, ex, ey, hz, _fict_)
