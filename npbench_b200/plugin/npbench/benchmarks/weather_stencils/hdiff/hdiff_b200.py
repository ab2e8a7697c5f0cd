# B200-native hdiff: same signature as hdiff_numpy.py:5 (bench_info/hdiff.json input_args).
from npbench_b200 import kernels as _k


def hdiff(in_field, out_field, coeff):
    _k.hdiff(in_field, out_field, coeff)
