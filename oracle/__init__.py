"""CPU oracle for the stencil hot path -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under npbench_b200/ may import this package.

ctypes front-end over oracle/liboracle.so (built from stencil_oracle.c by
oracle/build.py).  Every function mutates NumPy arrays in place exactly like
the NPBench NumPy function it restates (cited in stencil_oracle.c).
"""
import ctypes
import os

import numpy as np

from . import build as _build

_LIB = None
_dp = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64


def _ptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "oracle needs C-contiguous float64"
    return a.ctypes.data_as(_dp)


def lib():
    global _LIB
    if _LIB is None:
        path = _build.OUT
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(_build.SRC):
            path = _build.build()
        L = ctypes.CDLL(path)
        L.npb_oracle_set_threads.argtypes = [ctypes.c_int]
        L.npb_oracle_max_threads.restype = ctypes.c_int
        L.npb_oracle_jacobi2d.argtypes = [_i64, _i64, _i64, _dp, _dp]
        L.npb_oracle_jacobi2d_sweeps.argtypes = [_i64, _i64, _i64, _dp, _dp]
        L.npb_oracle_heat3d.argtypes = [_i64, _i64, _i64, _i64, _dp, _dp]
        L.npb_oracle_heat3d_sweeps.argtypes = [_i64, _i64, _i64, _i64, _dp, _dp]
        L.npb_oracle_fdtd2d.argtypes = [_i64, _i64, _i64, _dp, _dp, _dp, _dp]
        L.npb_oracle_hdiff.argtypes = [_i64, _i64, _i64, _dp, _dp, _dp]
        L.npb_oracle_vadv.argtypes = [_i64, _i64, _i64, _dp, _dp, _dp, _dp, _dp, ctypes.c_double]
        L.npb_oracle_init_jacobi2d.argtypes = [_i64, _i64, _i64, _i64, _dp, _dp]
        L.npb_oracle_init_heat3d.argtypes = [_i64, _i64, _i64, _dp, _dp]
        L.npb_oracle_init_fdtd2d.argtypes = [_i64, _i64, _i64, _i64, _i64, _dp, _dp, _dp, _dp]
        L.npb_oracle_jacobi1d.argtypes = [_i64, _i64, _dp, _dp]
        L.npb_oracle_init_jacobi1d.argtypes = [_i64, _dp, _dp]
        L.npb_oracle_seidel2d.argtypes = [_i64, _i64, _dp]
        L.npb_oracle_init_seidel2d.argtypes = [_i64, _dp]
        L.npb_oracle_adi.argtypes = [_i64, _i64, _dp]
        L.npb_oracle_init_adi.argtypes = [_i64, _dp]
        L.npb_oracle_cavity_flow.argtypes = [_i64, _i64, _i64, _i64, _dp, _dp, ctypes.c_double, ctypes.c_double,
                                             ctypes.c_double, _dp, ctypes.c_double, ctypes.c_double]
        L.npb_oracle_channel_flow.argtypes = [_i64, _i64, _i64, _dp, _dp, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                              _dp, ctypes.c_double, ctypes.c_double, ctypes.c_double, _i64]
        L.npb_oracle_channel_flow.restype = _i64
        L.npb_oracle_np_sum.argtypes = [_dp, _i64]
        L.npb_oracle_np_sum.restype = ctypes.c_double
        _LIB = L
    return _LIB


def set_threads(n: int) -> None:
    lib().npb_oracle_set_threads(int(n))


def max_threads() -> int:
    return int(lib().npb_oracle_max_threads())


# -- kernels (signatures mirror the NPBench NumPy functions) -----------------

def jacobi_2d(TSTEPS, A, B):
    ni, nj = A.shape
    assert B.shape == A.shape
    lib().npb_oracle_jacobi2d(int(TSTEPS), ni, nj, _ptr(A), _ptr(B))


def jacobi_2d_sweeps(nsweeps, A, B):
    ni, nj = A.shape
    lib().npb_oracle_jacobi2d_sweeps(int(nsweeps), ni, nj, _ptr(A), _ptr(B))


def heat_3d(TSTEPS, A, B):
    n0, n1, n2 = A.shape
    assert B.shape == A.shape
    lib().npb_oracle_heat3d(int(TSTEPS), n0, n1, n2, _ptr(A), _ptr(B))


def heat_3d_sweeps(nsweeps, A, B):
    n0, n1, n2 = A.shape
    lib().npb_oracle_heat3d_sweeps(int(nsweeps), n0, n1, n2, _ptr(A), _ptr(B))


def fdtd_2d(TMAX, ex, ey, hz, _fict_):
    nx, ny = ex.shape
    assert ey.shape == ex.shape and hz.shape == ex.shape and _fict_.shape[0] >= TMAX
    lib().npb_oracle_fdtd2d(int(TMAX), nx, ny, _ptr(ex), _ptr(ey), _ptr(hz), _ptr(_fict_))


def hdiff(in_field, out_field, coeff):
    I, J, K = out_field.shape
    assert in_field.shape == (I + 4, J + 4, K) and coeff.shape == (I, J, K)
    lib().npb_oracle_hdiff(I, J, K, _ptr(in_field), _ptr(out_field), _ptr(coeff))


def vadv(utens_stage, u_stage, wcon, u_pos, utens, dtr_stage):
    I, J, K = utens_stage.shape
    assert wcon.shape == (I + 1, J, K) and K >= 2
    lib().npb_oracle_vadv(I, J, K, _ptr(utens_stage), _ptr(u_stage), _ptr(wcon), _ptr(u_pos),
                          _ptr(utens), float(dtr_stage))


def jacobi_1d(TSTEPS, A, B):
    """polybench/jacobi_1d/jacobi_1d_numpy.py:4-8"""
    assert A.ndim == 1 and B.shape == A.shape
    lib().npb_oracle_jacobi1d(int(TSTEPS), A.shape[0], _ptr(A), _ptr(B))


def seidel_2d(TSTEPS, N, A):
    """polybench/seidel_2d/seidel_2d_numpy.py:4-13"""
    assert A.shape == (N, N)
    lib().npb_oracle_seidel2d(int(TSTEPS), int(N), _ptr(A))


def adi(TSTEPS, N, u):
    """polybench/adi/adi_numpy.py:6-54 (mutates u like the reference, which also returns it)"""
    assert u.shape == (N, N)
    lib().npb_oracle_adi(int(TSTEPS), int(N), _ptr(u))
    return u


def cavity_flow(nx, ny, nt, nit, u, v, dt, dx, dy, p, rho, nu):
    """cavity_flow/cavity_flow_numpy.py:46-89"""
    assert u.shape == (ny, nx) and v.shape == (ny, nx) and p.shape == (ny, nx) and nx >= 3 and ny >= 3
    lib().npb_oracle_cavity_flow(int(nx), int(ny), int(nt), int(nit), _ptr(u), _ptr(v), float(dt), float(dx), float(dy),
                                 _ptr(p), float(rho), float(nu))


def channel_flow(nit, u, v, dt, dx, dy, p, rho, nu, F, max_steps=10 ** 7):
    """channel_flow/channel_flow_numpy.py:74-170; returns stepcount like the reference"""
    ny, nx = u.shape
    assert v.shape == u.shape and p.shape == u.shape and nx >= 3 and ny >= 3
    return int(lib().npb_oracle_channel_flow(int(nit), nx, ny, _ptr(u), _ptr(v), float(dt), float(dx), float(dy), _ptr(p),
                                             float(rho), float(nu), float(F), int(max_steps)))


def np_sum(a):
    """NumPy's pairwise summation of a contiguous float64 array (np.sum), restated"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    return float(lib().npb_oracle_np_sum(_ptr(a), a.size))


# -- initialisers (NPBench's `initialize` functions restated) ----------------

def init_channel_flow(ny, nx):
    """channel_flow/channel_flow.py -> u, v, p, dx, dy, dt"""
    u = np.zeros((ny, nx)); v = np.zeros((ny, nx)); p = np.ones((ny, nx))
    return u, v, p, 2 / (nx - 1), 2 / (ny - 1), .1 / ((nx - 1) * (ny - 1))



def init_cavity_flow(ny, nx):
    """cavity_flow/cavity_flow.py:6-13 -> u, v, p, dx, dy, dt"""
    u = np.zeros((ny, nx)); v = np.zeros((ny, nx)); p = np.zeros((ny, nx))
    return u, v, p, 2 / (nx - 1), 2 / (ny - 1), .1 / ((nx - 1) * (ny - 1))



def init_adi(N):
    """adi.py: u = (i + N - j) / N"""
    u = np.empty((N, N))
    lib().npb_oracle_init_adi(N, _ptr(u))
    return u



def init_jacobi_1d(N):
    """jacobi_1d.py:6-10"""
    A = np.empty((N,)); B = np.empty((N,))
    lib().npb_oracle_init_jacobi1d(N, _ptr(A), _ptr(B))
    return A, B


def init_seidel_2d(N):
    """seidel_2d.py:6-10"""
    A = np.empty((N, N))
    lib().npb_oracle_init_seidel2d(N, _ptr(A))
    return A



def init_jacobi_2d(N, row0=0, nrows=None, ncols=None):
    """jacobi_2d.py:6-10; optional slab [row0, row0+nrows) x ncols of an N-row grid."""
    nrows = N if nrows is None else nrows
    ncols = N if ncols is None else ncols
    A = np.empty((nrows, ncols)); B = np.empty((nrows, ncols))
    lib().npb_oracle_init_jacobi2d(N, row0, nrows, ncols, _ptr(A), _ptr(B))
    return A, B


def init_heat_3d(N, row0=0, nrows=None):
    """heat_3d.py:6-11."""
    nrows = N if nrows is None else nrows
    A = np.empty((nrows, N, N)); B = np.empty((nrows, N, N))
    lib().npb_oracle_init_heat3d(N, row0, nrows, _ptr(A), _ptr(B))
    return A, B


def init_fdtd_2d(TMAX, NX, NY, row0=0, nrows=None):
    """fdtd_2d.py:6-15."""
    nrows = NX if nrows is None else nrows
    ex = np.empty((nrows, NY)); ey = np.empty((nrows, NY)); hz = np.empty((nrows, NY))
    fict = np.empty((TMAX,))
    lib().npb_oracle_init_fdtd2d(TMAX, NX, NY, row0, nrows, _ptr(ex), _ptr(ey), _ptr(hz), _ptr(fict))
    return ex, ey, hz, fict


def init_hdiff(I, J, K):
    """hdiff.py:6-15 -- default_rng(42) draw order: in_field, out_field, coeff."""
    rng = np.random.default_rng(42)
    in_field = rng.random((I + 4, J + 4, K))
    out_field = rng.random((I, J, K))
    coeff = rng.random((I, J, K))
    return in_field, out_field, coeff


def init_vadv(I, J, K):
    """vadv.py:6-19 -- default_rng(42) draw order: utens_stage, u_stage, wcon, u_pos, utens."""
    rng = np.random.default_rng(42)
    dtr_stage = 3.0 / 20.0
    utens_stage = rng.random((I, J, K))
    u_stage = rng.random((I, J, K))
    wcon = rng.random((I + 1, J, K))
    u_pos = rng.random((I, J, K))
    utens = rng.random((I, J, K))
    return dtr_stage, utens_stage, u_stage, wcon, u_pos, utens


# NPBench presets: bench_info/{jacobi_2d,heat_3d,fdtd_2d,hdiff,vadv}.json:11-16
PRESETS = {
    "jacobi_2d": {"S": dict(TSTEPS=50, N=150), "M": dict(TSTEPS=80, N=350),
                  "L": dict(TSTEPS=200, N=700), "paper": dict(TSTEPS=1000, N=2800)},
    "heat_3d": {"S": dict(TSTEPS=25, N=25), "M": dict(TSTEPS=50, N=40),
                "L": dict(TSTEPS=100, N=70), "paper": dict(TSTEPS=500, N=120)},
    "fdtd_2d": {"S": dict(TMAX=20, NX=200, NY=220), "M": dict(TMAX=60, NX=400, NY=450),
                "L": dict(TMAX=150, NX=800, NY=900), "paper": dict(TMAX=500, NX=1000, NY=1200)},
    "hdiff": {"S": dict(I=64, J=64, K=60), "M": dict(I=128, J=128, K=160),
              "L": dict(I=384, J=384, K=160), "paper": dict(I=256, J=256, K=160)},
    "vadv": {"S": dict(I=60, J=60, K=40), "M": dict(I=112, J=112, K=80),
             "L": dict(I=180, J=180, K=160), "paper": dict(I=256, J=256, K=160)},
    # widening row: bench_info/{jacobi_1d,seidel_2d}.json:11-16
    "jacobi_1d": {"S": dict(TSTEPS=800, N=3200), "M": dict(TSTEPS=3000, N=12000),
                  "L": dict(TSTEPS=8500, N=34000), "paper": dict(TSTEPS=4000, N=32000)},
    "seidel_2d": {"S": dict(TSTEPS=8, N=50), "M": dict(TSTEPS=15, N=100),
                  "L": dict(TSTEPS=40, N=200), "paper": dict(TSTEPS=100, N=400)},
    "cavity_flow": {"S": dict(ny=61, nx=61, nt=25, nit=5, rho=1.0, nu=0.1), "M": dict(ny=121, nx=121, nt=50, nit=10, rho=1.0, nu=0.1),
                    "L": dict(ny=201, nx=201, nt=100, nit=20, rho=1.0, nu=0.1),
                    "paper": dict(ny=101, nx=101, nt=700, nit=50, rho=1.0, nu=0.1)},
    "channel_flow": {"S": dict(ny=61, nx=61, nit=5, rho=1.0, nu=0.1, F=1.0), "M": dict(ny=121, nx=121, nit=10, rho=1.0, nu=0.1, F=1.0),
                     "L": dict(ny=201, nx=201, nit=20, rho=1.0, nu=0.1, F=1.0),
                     "paper": dict(ny=101, nx=101, nit=50, rho=1.0, nu=0.1, F=1.0)},
    "adi": {"S": dict(TSTEPS=5, N=100), "M": dict(TSTEPS=20, N=200),
            "L": dict(TSTEPS=50, N=500), "paper": dict(TSTEPS=100, N=200)},
}
