# B200-native jacobi_1d: same signature as jacobi_1d_numpy.py:4 (bench_info/jacobi_1d.json input_args).
from npbench_b200 import kernels as _k


def kernel(TSTEPS, A, B):
    _k.jacobi_1d(TSTEPS, A, B)
