"""Host-side logic of the slab-sharded drivers (npbench_b200/distributed.py) on CPU.

world_size 2 and 3, gloo backend, 127.0.0.1.  The CUDA engine is replaced by a CPU engine
that computes each local operation with the oracle and POISONS (NaN) everything the CUDA
kernels leave undefined on ghost rows, so any leak of ghost garbage into owned rows, a
missing exchange or a wrong tile/row range shows up as a bit mismatch against the
single-domain oracle result.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from npbench_b200 import distributed as D


class CpuEngine:
    tile_rows = 4          # small tiles => many boundary/interior splits

    def empty(self, *shape):
        return torch.full(shape, float("nan"), dtype=torch.float64)

    dual = False           # True: the slab driver may use passes that store two states (marching regime on the GPU)

    def jacobi_dual_ok(self, nrows, ncols):
        return self.dual

    def jacobi_block(self, n, src, dst, t_lo, t_hi, dst2=None):
        a, b = src.numpy().copy(), dst.numpy().copy()
        r0, r1 = 1 + t_lo * self.tile_rows, min(src.shape[0] - 1, 1 + t_hi * self.tile_rows)
        if dst2 is not None:              # state n - 1 (even number of sweeps: it ends in `a`) -> dst2, then the last sweep
            assert n % 2 == 1 and n >= 3
            oracle.jacobi_2d_sweeps(n - 1, a, b)
            dst2.numpy()[r0:r1, 1:-1] = a[r0:r1, 1:-1]
            oracle.jacobi_2d_sweeps(1, a, b)
            res = b
        else:
            oracle.jacobi_2d_sweeps(n, a, b)
            res = b if n % 2 else a
        dst.numpy()[r0:r1, 1:-1] = res[r0:r1, 1:-1]

    def heat_sweep(self, src, dst, i_lo, i_hi):
        a, b = src.numpy().copy(), dst.numpy().copy()
        oracle.heat_3d_sweeps(1, a, b)
        i_lo, i_hi = max(1, i_lo), min(src.shape[0] - 1, i_hi)
        dst.numpy()[i_lo:i_hi, 1:-1, 1:-1] = b[i_lo:i_hi, 1:-1, 1:-1]

    def heat_march(self, src, dst, i_lo, i_hi):
        """three sweeps src -> dst like heat3d_march_kernel: planes 0 / n-1 act as constant borders (state 1 takes them
        from dst, state 2 from src); on a slab they are ghost planes, so everything within 3 planes of a slab edge that
        is not a grid edge is garbage in the CUDA contract: poisoned here."""
        a, b = src.numpy().copy(), dst.numpy().copy()
        oracle.heat_3d_sweeps(3, a, b)
        n = src.shape[0]
        slab = self.slab
        if slab.ht:
            b[1:3] = np.nan
        if slab.hb:
            b[n - 3:n - 1] = np.nan
        i_lo, i_hi = max(1, i_lo), min(n - 1, i_hi)
        dst.numpy()[i_lo:i_hi, 1:-1, 1:-1] = b[i_lo:i_hi, 1:-1, 1:-1]

    def fict_on_device(self, fict):
        return np.asarray(fict, dtype=np.float64)

    def fdtd_march(self, ns, nx_global, row0, src, dst, fict_dev, t, r_lo, r_hi):
        """ns steps src -> dst like fdtd2d_march_kernel: rows within ns of a slab edge that is not a grid edge are garbage
        in the CUDA contract (ghost rows): poisoned here."""
        assert 2 <= ns <= 5
        f = [x.numpy().copy() for x in src]
        n = f[0].shape[0]
        oracle.fdtd_2d(ns, f[0], f[1], f[2], np.ascontiguousarray(fict_dev[t:t + ns]))
        if row0 > 0:
            for x in f:
                x[:ns] = np.nan
        if row0 + n < nx_global:
            for x in f:
                x[n - ns:] = np.nan
        r_hi = n if r_hi < 0 else r_hi
        for d, x in zip(dst, f):
            d.numpy()[r_lo:r_hi] = x[r_lo:r_hi]

    def fdtd_step(self, nx_global, row0, src, dst, fict_t, r_lo, r_hi):
        f = [x.numpy().copy() for x in src]
        n = f[0].shape[0]
        oracle.fdtd_2d(1, f[0], f[1], f[2], np.array([fict_t]))
        if row0 > 0:                      # ghost row: no row above -> undefined in the CUDA contract
            f[1][0] = np.nan; f[2][0] = np.nan
        if row0 + n < nx_global:          # ghost row: no row below
            f[2][n - 1] = np.nan
        r_hi = n if r_hi < 0 else r_hi
        for d, x in zip(dst, f):
            d.numpy()[r_lo:r_hi] = x[r_lo:r_hi]

    def copy(self, dst, src):
        dst.copy_(src)

    def boundary_done(self):
        return None

    def start_exchange(self, exchanger, fields, after):
        return exchanger.start(fields)

    def finish_exchange(self, reqs):
        D.HaloExchanger.finish(reqs)


def _local(slab, full):
    return torch.from_numpy(np.ascontiguousarray(full[slab.row0:slab.row0 + slab.nloc]).copy())


def _worker(rank, size, port, case):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        eng = CpuEngine()
        rng = np.random.default_rng(1234)          # same stream on every rank
        if case == "jacobi":
            for ts, (ni, nj) in ((2, (40, 13)), (6, (61, 30)), (12, (47, 19)), (3, (44, 9))):
                A0, B0 = rng.random((ni, nj)), rng.random((ni, nj))
                for dual in (False, True):      # True: the closing pass stores the last two states (scratch slab W)
                    A, B = A0.copy(), B0.copy()
                    eng.dual = dual
                    slab = D.Slab(ni, size, rank, D.JACOBI_MAX_BLOCK)
                    lA, lB = _local(slab, A), _local(slab, B)
                    D.jacobi_2d_sharded(eng, slab, ts, lA, lB)
                    oracle.jacobi_2d(ts, A, B)
                    assert np.array_equal(slab.owned(lA).numpy(), A[slab.lo:slab.hi]), ("A", ts, rank, dual)
                    assert np.array_equal(slab.owned(lB).numpy(), B[slab.lo:slab.hi]), ("B", ts, rank, dual)
        elif case == "heat":
            for ts, shape, H in ((2, (14, 6, 7), 3), (5, (17, 5, 6), 4), (8, (25, 6, 5), 3), (7, (40, 5, 6), 3), (4, (36, 4, 5), 5)):
                A, B = rng.random(shape), rng.random(shape)
                for march in (False, True):
                    a, b = A.copy(), B.copy()
                    slab = D.Slab(shape[0], size, rank, H)
                    eng.slab = slab
                    lA, lB = _local(slab, a), _local(slab, b)
                    D.heat_3d_sharded(eng, slab, ts, lA, lB, march=march and shape[0] // size + H >= 8)
                    oracle.heat_3d(ts, a, b)
                    assert np.array_equal(slab.owned(lA).numpy(), a[slab.lo:slab.hi]), ("A", ts, rank, march)
                    assert np.array_equal(slab.owned(lB).numpy(), b[slab.lo:slab.hi]), ("B", ts, rank, march)
        elif case == "fdtd":
            for tm, (nx, ny), H in ((1, (12, 9), 2), (7, (23, 11), 3), (10, (30, 8), 4), (4, (19, 5), 5), (13, (41, 6), 5), (11, (33, 7), 7)):
                ex, ey, hz = rng.random((nx, ny)), rng.random((nx, ny)), rng.random((nx, ny))
                fict = rng.random(tm)
                for march in (False, True):
                    g = [f.copy() for f in (ex, ey, hz)]
                    slab = D.Slab(nx, size, rank, H)
                    eng.slab = slab
                    l = [_local(slab, f) for f in g]
                    D.fdtd_2d_sharded(eng, slab, tm, l[0], l[1], l[2], fict, march=march)
                    oracle.fdtd_2d(tm, g[0], g[1], g[2], fict)
                    for name, got, want in zip(("ex", "ey", "hz"), l, g):
                        assert np.array_equal(slab.owned(got).numpy(), want[slab.lo:slab.hi]), (name, tm, rank, march)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("size", [2, 3])
@pytest.mark.parametrize("case", ["jacobi", "heat", "fdtd"])
def test_sharded_equals_single_domain(case, size):
    mp.spawn(_worker, args=(size, _free_port(), case), nprocs=size, join=True)


def test_single_rank_driver_equals_kernel():
    """size == 1: no process group needed, the driver degenerates to the plain kernel."""
    eng = CpuEngine()
    rng = np.random.default_rng(5)
    A, B = rng.random((21, 17)), rng.random((21, 17))
    slab = D.Slab(21, 1, 0, D.JACOBI_MAX_BLOCK)
    lA, lB = torch.from_numpy(A.copy()), torch.from_numpy(B.copy())
    D.jacobi_2d_sharded(eng, slab, 9, lA, lB)
    oracle.jacobi_2d(9, A, B)
    assert np.array_equal(lA.numpy(), A) and np.array_equal(lB.numpy(), B)
    eng.dual = True                    # the scratch-slab plan on one rank
    A2, B2 = rng.random((21, 17)), rng.random((21, 17))
    lA, lB = torch.from_numpy(A2.copy()), torch.from_numpy(B2.copy())
    D.jacobi_2d_sharded(eng, slab, 9, lA, lB)
    oracle.jacobi_2d(9, A2, B2)
    assert np.array_equal(lA.numpy(), A2) and np.array_equal(lB.numpy(), B2)


def test_plan_matches_the_c_driver_rules():
    for ts in range(1, 60):
        plan = D.jacobi_plan(2 * (ts - 1))
        assert sum(plan) == 2 * (ts - 1)
        assert all(p % 2 == 1 and 1 <= p <= 7 for p in plan)
        assert len(plan) % 2 == 0 and (not plan or plan[-1] == 1)


def test_dual_plan_rules():
    """an even number of odd passes of at most 7 sweeps, ascending, that cover all sweeps (csrc/jacobi2d.cu: the
    scratch-grid plan of npb_jacobi2d_f64)"""
    for ts in range(2, 60):
        plan = D.jacobi_plan_dual(2 * (ts - 1))
        assert sum(plan) == 2 * (ts - 1) and len(plan) % 2 == 0
        assert all(p % 2 == 1 and 1 <= p <= 7 for p in plan) and plan == sorted(plan)
    assert D.jacobi_plan_dual(40) == [5, 7, 7, 7, 7, 7]


def test_slab_bounds_cover_without_overlap():
    for n in (7, 64, 1000):
        for size in (1, 2, 3, 8):
            b = [D.slab_bounds(n, size, r) for r in range(size)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(size - 1))


@pytest.mark.parametrize("size", [2, 4])
def test_halo_free_shards_of_hdiff_and_vadv(size):
    I, J, K = 13, 6, 5
    inf, outf, coeff = oracle.init_hdiff(I, J, K)
    want = outf.copy(); oracle.hdiff(inf, want, coeff)
    parts = []
    for r in range(size):
        lo, hi = D.hdiff_shard(I, size, r)
        o = np.empty((hi - lo, J, K))
        oracle.hdiff(np.ascontiguousarray(inf[lo:hi + 4]), o, np.ascontiguousarray(coeff[lo:hi]))
        parts.append(o)
    assert np.array_equal(np.concatenate(parts), want)
    dtr, us, u, w, up, ut = oracle.init_vadv(I, J, K)
    want = us.copy(); oracle.vadv(want, u, w, up, ut, dtr)
    parts = []
    for r in range(size):
        lo, hi = D.vadv_shard(I, size, r)
        o = np.ascontiguousarray(us[lo:hi]).copy()
        oracle.vadv(o, *(np.ascontiguousarray(x[lo:hi]) for x in (u,)), np.ascontiguousarray(w[lo:hi + 1]),
                    np.ascontiguousarray(up[lo:hi]), np.ascontiguousarray(ut[lo:hi]), dtr)
        parts.append(o)
    assert np.array_equal(np.concatenate(parts), want)
