# B200-native adi: same signature and return value as adi_numpy.py:6-54 (bench_info/adi.json input_args).
from npbench_b200 import kernels as _k


def kernel(TSTEPS, N, u):
    return _k.adi(TSTEPS, N, u)
