# B200-native jacobi_2d: same signature as jacobi_2d_numpy.py:4 (bench_info/jacobi_2d.json input_args).
from npbench_b200 import kernels as _k


def kernel(TSTEPS, A, B):
    _k.jacobi_2d(TSTEPS, A, B)
