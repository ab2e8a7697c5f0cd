# B200-native cavity_flow: same signature as cavity_flow_numpy.py:46 (bench_info/cavity_flow.json input_args).
from npbench_b200 import kernels as _k


def cavity_flow(nx, ny, nt, nit, u, v, dt, dx, dy, p, rho, nu):
    _k.cavity_flow(nx, ny, nt, nit, u, v, dt, dx, dy, p, rho, nu)
