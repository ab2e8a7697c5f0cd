# B200-native seidel_2d: same signature as seidel_2d_numpy.py:4 (bench_info/seidel_2d.json input_args).
from npbench_b200 import kernels as _k


def kernel(TSTEPS, N, A):
    _k.seidel_2d(TSTEPS, N, A)
