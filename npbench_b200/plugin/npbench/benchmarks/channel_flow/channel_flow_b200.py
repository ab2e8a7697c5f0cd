# B200-native channel_flow: same signature and return value (stepcount) as channel_flow_numpy.py:74-170
# (bench_info/channel_flow.json input_args).
from npbench_b200 import kernels as _k


def channel_flow(nit, u, v, dt, dx, dy, p, rho, nu, F):
    return _k.channel_flow(nit, u, v, dt, dx, dy, p, rho, nu, F)
