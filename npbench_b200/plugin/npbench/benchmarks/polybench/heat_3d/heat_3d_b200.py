# B200-native heat_3d: same signature as heat_3d_numpy.py:4 (bench_info/heat_3d.json input_args).
from npbench_b200 import kernels as _k


def kernel(TSTEPS, A, B):
    _k.heat_3d(TSTEPS, A, B)
