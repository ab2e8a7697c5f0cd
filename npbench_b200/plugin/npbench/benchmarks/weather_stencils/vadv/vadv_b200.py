# B200-native vadv: same signature as vadv_numpy.py:9 (bench_info/vadv.json input_args).
from npbench_b200 import kernels as _k


def vadv(utens_stage, u_stage, wcon, u_pos, utens, dtr_stage):
    _k.vadv(utens_stage, u_stage, wcon, u_pos, utens, dtr_stage)
