"""Host-side mirror of the five NPBench functions, with their exact NumPy signatures.

Each function accepts either DeviceArray arguments (what B200Framework.copy_func
produced: the kernel is enqueued on the library stream and the call returns --
the harness's exec_str appends the sync) or NumPy arrays (the C-ABI *_host entry
point copies in, runs, copies the outputs back and synchronises).  Like the
reference functions they mutate their array arguments and return None.

Reference signatures: bench_info/{jacobi_2d,heat_3d,fdtd_2d,hdiff,vadv}.json
"input_args"; implementations npbench/benchmarks/**/<bench>_numpy.py.
"""
import ctypes

import numpy as np

from . import _lib
from .device_array import DeviceArray


def _kind(*arrays):
    dev = [isinstance(a, DeviceArray) for a in arrays]
    if all(dev):
        return "device"
    if not any(dev):
        for a in arrays:
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]):
                raise TypeError("array arguments must be C-contiguous float64 (NPBench's array_args are)")
        return "host"
    raise TypeError("mixing DeviceArray and numpy arguments is not supported")


def _p(a):
    return a.ptr if isinstance(a, DeviceArray) else a.ctypes.data


def jacobi_2d(TSTEPS, A, B):
    """kernel(TSTEPS, A, B) -- polybench/jacobi_2d/jacobi_2d_numpy.py:4-10."""
    if A.shape != B.shape or len(A.shape) != 2:
        raise ValueError("A and B must be 2-D arrays of the same shape")
    L = _lib.lib()
    fn = L.jacobi2d_f64 if _kind(A, B) == "device" else L.jacobi2d_f64_host
    fn(int(TSTEPS), A.shape[0], A.shape[1], _p(A), _p(B))


def heat_3d(TSTEPS, A, B):
    """kernel(TSTEPS, A, B) -- polybench/heat_3d/heat_3d_numpy.py:4-20."""
    if A.shape != B.shape or len(A.shape) != 3:
        raise ValueError("A and B must be 3-D arrays of the same shape")
    L = _lib.lib()
    fn = L.heat3d_f64 if _kind(A, B) == "device" else L.heat3d_f64_host
    fn(int(TSTEPS), A.shape[0], A.shape[1], A.shape[2], _p(A), _p(B))


def fdtd_2d(TMAX, ex, ey, hz, _fict_):
    """kernel(TMAX, ex, ey, hz, _fict_) -- polybench/fdtd_2d/fdtd_2d_numpy.py:4-11."""
    if not (ex.shape == ey.shape == hz.shape) or len(ex.shape) != 2:
        raise ValueError("ex, ey, hz must be 2-D arrays of the same shape")
    if _fict_.shape[0] < TMAX:
        raise IndexError("_fict_ has fewer than TMAX entries")   # NumPy would raise at _fict_[t]
    L = _lib.lib()
    fn = L.fdtd2d_f64 if _kind(ex, ey, hz, _fict_) == "device" else L.fdtd2d_f64_host
    fn(int(TMAX), ex.shape[0], ex.shape[1], _p(ex), _p(ey), _p(hz), _p(_fict_))


def hdiff(in_field, out_field, coeff):
    """hdiff(in_field, out_field, coeff) -- weather_stencils/hdiff/hdiff_numpy.py:5-29."""
    from .sharded import ShardedArray, hdiff_mg
    if isinstance(out_field, ShardedArray):            # NPB_B200_GPUS > 1: column shards, one process, no exchange
        return hdiff_mg(in_field, out_field, coeff)
    I, J, K = out_field.shape
    if tuple(in_field.shape) != (I + 4, J + 4, K) or tuple(coeff.shape) != (I, J, K):
        raise ValueError("expected in_field (I+4,J+4,K), out_field/coeff (I,J,K)")
    L = _lib.lib()
    fn = L.hdiff_f64 if _kind(in_field, out_field, coeff) == "device" else L.hdiff_f64_host
    fn(I, J, K, _p(in_field), _p(out_field), _p(coeff))


def vadv(utens_stage, u_stage, wcon, u_pos, utens, dtr_stage):
    """vadv(utens_stage, u_stage, wcon, u_pos, utens, dtr_stage) -- vadv_numpy.py:9-78."""
    from .sharded import ShardedArray, vadv_mg
    if isinstance(utens_stage, ShardedArray):
        return vadv_mg(utens_stage, u_stage, wcon, u_pos, utens, dtr_stage)
    I, J, K = utens_stage.shape
    if tuple(wcon.shape) != (I + 1, J, K):
        raise ValueError("wcon must be (I+1, J, K)")
    for a in (u_stage, u_pos, utens):
        if tuple(a.shape) != (I, J, K):
            raise ValueError("u_stage, u_pos, utens must match utens_stage")
    if K < 2:
        raise IndexError("vadv needs K >= 2 (the reference indexes level k+1 at k = 0)")
    L = _lib.lib()
    fn = L.vadv_f64 if _kind(utens_stage, u_stage, wcon, u_pos, utens) == "device" else L.vadv_f64_host
    fn(I, J, K, _p(utens_stage), _p(u_stage), _p(wcon), _p(u_pos), _p(utens), float(dtr_stage))


def sync():
    """Block until everything enqueued on the library stream has finished."""
    _lib.lib().sync()


# ---- widening row: bench_info/{jacobi_1d,seidel_2d}.json input_args ----

def jacobi_1d(TSTEPS, A, B):
    """kernel(TSTEPS, A, B) -- polybench/jacobi_1d/jacobi_1d_numpy.py:4-8."""
    if A.shape != B.shape or len(A.shape) != 1:
        raise ValueError("A and B must be 1-D arrays of the same length")
    L = _lib.lib()
    fn = L.jacobi1d_f64 if _kind(A, B) == "device" else L.jacobi1d_f64_host
    fn(int(TSTEPS), A.shape[0], _p(A), _p(B))


def seidel_2d(TSTEPS, N, A):
    """kernel(TSTEPS, N, A) -- polybench/seidel_2d/seidel_2d_numpy.py:4-13."""
    if len(A.shape) != 2 or A.shape[0] != A.shape[1] or A.shape[0] != int(N):
        raise ValueError("A must be an (N, N) array")
    L = _lib.lib()
    fn = L.seidel2d_f64 if _kind(A) == "device" else L.seidel2d_f64_host
    fn(int(TSTEPS), int(N), _p(A))


def adi(TSTEPS, N, u):
    """kernel(TSTEPS, N, u) -- polybench/adi/adi_numpy.py:6-54; returns u like the reference (:54)."""
    if len(u.shape) != 2 or u.shape[0] != u.shape[1] or u.shape[0] != int(N):
        raise ValueError("u must be an (N, N) array")
    if int(TSTEPS) == 0:
        raise ZeroDivisionError("float division by zero")       # DT = 1.0 / TSTEPS, adi_numpy.py:14
    L = _lib.lib()
    if int(TSTEPS) > 0:
        fn = L.adi_f64 if _kind(u) == "device" else L.adi_f64_host
        fn(int(TSTEPS), int(N), _p(u))
    return u


def cavity_flow(nx, ny, nt, nit, u, v, dt, dx, dy, p, rho, nu):
    """cavity_flow(nx, ny, nt, nit, u, v, dt, dx, dy, p, rho, nu) -- cavity_flow/cavity_flow_numpy.py:46-89."""
    shape = (int(ny), int(nx))
    if tuple(u.shape) != shape or tuple(v.shape) != shape or tuple(p.shape) != shape:
        raise ValueError("u, v and p must be (ny, nx) arrays")
    if int(nx) < 3 or int(ny) < 3:
        raise ValueError("nx and ny must be >= 3")
    L = _lib.lib()
    fn = L.cavity_flow_f64 if _kind(u, v, p) == "device" else L.cavity_flow_f64_host
    fn(int(nx), int(ny), int(nt), int(nit), _p(u), _p(v), float(dt), float(dx), float(dy), _p(p), float(rho), float(nu))


def channel_flow(nit, u, v, dt, dx, dy, p, rho, nu, F):
    """channel_flow(nit, u, v, dt, dx, dy, p, rho, nu, F) -> stepcount -- channel_flow/channel_flow_numpy.py:74-170."""
    if len(u.shape) != 2 or tuple(v.shape) != tuple(u.shape) or tuple(p.shape) != tuple(u.shape):
        raise ValueError("u, v and p must be 2-D arrays of the same shape")
    ny, nx = u.shape
    if nx < 3 or ny < 3:
        raise ValueError("nx and ny must be >= 3")
    L = _lib.lib()
    fn = L.channel_flow_f64 if _kind(u, v, p) == "device" else L.channel_flow_f64_host
    steps = ctypes.c_int64(0)
    fn(int(nit), nx, ny, _p(u), _p(v), float(dt), float(dx), float(dy), _p(p), float(rho), float(nu), float(F),
       ctypes.addressof(steps))
    return int(steps.value)
