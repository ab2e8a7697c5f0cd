"""Host logic of the register-tile resident kernels (jacobi2d_regtile.cuh, fdtd2d_regtile.cuh), on the CPU.

Two things are checked without a GPU:
  * the configuration the library picks for a grid (npb_jacobi2d_regtile_plan / npb_fdtd2d_regtile_plan: pure host
    functions of the C ABI) satisfies every constraint the kernels rely on;
  * the SCHEME itself -- tiles with T-deep redundant halos, the whole region updated every sweep, garbage creeping in
    from the region edges, tile cells within T of an edge sent to the <= 3 neighbours that hold them as halo, the
    border ring re-imposed per state parity -- reproduces the oracle bit for bit.  The emulation below restates the
    kernels' geometry code (tile_bounds / cell_range, the role masks, the send targets) in NumPy and asserts on the
    way that every halo cell has exactly one sender and that no garbage reaches a tile.
The GPU tests (tests/test_parity_gpu.py::test_*_regtile_forced_configurations) run the real kernels.
"""
import ctypes

import numpy as np
import pytest

import oracle
from npbench_b200 import _lib


def plan_j2(ts, ni, nj, sms):
    out = (ctypes.c_int * 7)()
    ok = _lib.lib().jacobi2d_regtile_plan(ts, ni, nj, sms, ctypes.cast(out, ctypes.c_void_p))
    return list(out) if ok == 1 else None


def plan_f2(tm, nx, ny, sms):
    out = (ctypes.c_int * 6)()
    ok = _lib.lib().fdtd2d_regtile_plan(tm, nx, ny, sms, ctypes.cast(out, ctypes.c_void_p))
    return list(out) if ok == 1 else None


def split(n, parts, t, first):
    """tile_bounds (first = 1: interior cells) / cell_range (first = 0): [lo, hi) of part t"""
    base, rem = divmod(n, parts)
    lo = first + t * base + min(t, rem)
    return lo, lo + base + (1 if t < rem else 0)


def check_plan(n0, n1, sms, rb, cb, nw, T, PI, PJ, per_sm=1):
    assert rb in (2, 4, 8) and cb in (2, 4) and 1 <= nw and nw * 32 * per_sm <= 1024 and T >= 1 and per_sm in (1, 2)
    assert PI * PJ <= sms * per_sm
    for n, parts, cap, blk in ((n0, PI, nw * rb - 2 * T, rb), (n1, PJ, 32 * cb - 2 * T, cb)):
        assert cap >= 1 and -(-n // parts) <= cap                   # the largest tile + 2 T fits the region
        if parts > 1:
            assert n // parts >= 2 * T + blk - 1                    # halos from adjacent tiles only; <= 3 send targets


@pytest.mark.parametrize("sms", [148, 132, 20, 4])
def test_jacobi2d_plan_constraints(sms):
    seen = 0
    for ts, n in [(50, 150), (80, 350), (200, 700), (2, 40), (3, 33), (10, 1000), (30, 97), (5, 513), (1000, 2800)]:
        for nj in (n, n + 37, max(3, n // 3)):
            p = plan_j2(ts, n, nj, sms)
            if p is None:
                continue
            seen += 1
            check_plan(n - 2, nj - 2, sms, *p)
    assert seen >= 6
    assert plan_j2(1000, 2800, 2800, 148) is None                  # `paper` does not fit on chip
    assert plan_j2(1, 100, 100, 148) is None                       # no sweeps


@pytest.mark.parametrize("sms", [148, 20])
def test_fdtd2d_plan_constraints(sms):
    seen = 0
    for tm, nx, ny in [(20, 200, 220), (60, 400, 450), (150, 800, 900), (5, 30, 1), (7, 1, 500), (3, 64, 64), (9, 257, 339)]:
        p = plan_f2(tm, nx, ny, sms)
        if p is None:
            continue
        seen += 1
        check_plan(nx, ny, sms, *p)
    assert seen >= 3
    assert plan_f2(500, 1000, 1200, 148) is None                   # `paper` does not fit on chip


def test_presets_run_resident_on_a_b200():
    assert plan_j2(50, 150, 150, 148) and plan_j2(80, 350, 350, 148) and plan_j2(200, 700, 700, 148)
    assert plan_f2(20, 200, 220, 148) and plan_f2(60, 400, 450, 148) and plan_f2(150, 800, 900, 148)


# ------------------------------------------------------------------ the scheme, emulated
GARBAGE = 12345.678


class Tiles:
    """geometry shared by both emulations: region (r, c) <-> global (lo_i - T + r, lo_j - T + c)"""

    def __init__(self, n0, n1, first, rb, cb, nw, T, PI, PJ):
        self.RR, self.RC, self.T, self.PI, self.PJ = nw * rb, 32 * cb, T, PI, PJ
        self.t = {}
        for ti in range(PI):
            for tj in range(PJ):
                ilo, ihi = split(n0, PI, ti, first)
                jlo, jhi = split(n1, PJ, tj, first)
                assert ihi - ilo + 2 * T <= self.RR and jhi - jlo + 2 * T <= self.RC
                gi = (ilo - T + np.arange(self.RR))[:, None] + np.zeros((1, self.RC), int)
                gj = (jlo - T + np.arange(self.RC))[None, :] + np.zeros((self.RR, 1), int)
                own = (gi >= ilo) & (gi < ihi) & (gj >= jlo) & (gj < jhi)
                need = (gi >= ilo - T) & (gi < ihi + T) & (gj >= jlo - T) & (gj < jhi + T)
                self.t[ti, tj] = dict(ilo=ilo, ihi=ihi, jlo=jlo, jhi=jhi, gi=gi, gj=gj, own=own, need=need, inbox={})

    def exchange(self, fields):
        """fields: name -> {tile key -> region array of the NEW state}; sends exactly as the kernels' Desc.off / msk do"""
        T, PI, PJ = self.T, self.PI, self.PJ
        for (ti, tj), t in self.t.items():
            gi, gj = t["gi"], t["gj"]
            r_own = (gi[:, 0] >= t["ilo"]) & (gi[:, 0] < t["ihi"])
            r_top = r_own & (gi[:, 0] - t["ilo"] < T) & (ti > 0)
            r_bot = r_own & (gi[:, 0] >= t["ihi"] - T) & (ti < PI - 1)
            c_own = (gj[0] >= t["jlo"]) & (gj[0] < t["jhi"])
            c_lft = c_own & (gj[0] - t["jlo"] < T) & (tj > 0)
            c_rgt = c_own & (gj[0] >= t["jhi"] - T) & (tj < PJ - 1)
            for di, rs in ((-1, r_top), (0, r_own), (1, r_bot)):
                for dj, cs in ((-1, c_lft), (0, c_own), (1, c_rgt)):
                    if (di == 0 and dj == 0) or not rs.any() or not cs.any():
                        continue
                    n = self.t[ti + di, tj + dj]
                    for r in np.nonzero(rs)[0]:
                        for c in np.nonzero(cs)[0]:
                            rn, cn = gi[r, c] - n["ilo"] + T, gj[r, c] - n["jlo"] + T
                            assert 0 <= rn < self.RR and 0 <= cn < self.RC and n["halo"][rn, cn]
                            assert (rn, cn) not in n["inbox"], "two senders for one halo cell"
                            n["inbox"][rn, cn] = {k: v[ti, tj][r, c] for k, v in fields.items()}
        for key, t in self.t.items():
            hs = np.argwhere(t["halo"])
            assert len(hs) == len(t["inbox"]), "a halo cell without sender"
            for r, c in hs:
                for k, v in fields.items():
                    v[key][r, c] = t["inbox"][r, c][k]
            t["inbox"] = {}


def emulate_jacobi(A, B, tsteps, rb, cb, nw, T, PI, PJ, per_sm=1):
    ni, nj = A.shape
    nsweeps = 2 * (tsteps - 1)
    g = Tiles(ni - 2, nj - 2, 1, rb, cb, nw, T, PI, PJ)
    o, fx = {}, {}
    for key, t in g.t.items():
        gi, gj = t["gi"], t["gj"]
        inside = (gi >= 0) & (gi < ni) & (gj >= 0) & (gj < nj)
        interior = (gi >= 1) & (gi <= ni - 2) & (gj >= 1) & (gj <= nj - 2)
        t["interior"] = interior
        t["halo"] = interior & ~t["own"] & t["need"]
        ci, cj = np.clip(gi, 0, ni - 1), np.clip(gj, 0, nj - 1)
        fx[key] = (np.where(inside & ~interior, A[ci, cj], 0.0), np.where(inside & ~interior, B[ci, cj], 0.0))
        o[key] = np.where(inside, A[ci, cj], 0.0)
    outA, outB = A.copy(), B.copy()
    for s in range(1, nsweeps + 1):
        v = {}
        for key in g.t:
            p = np.full((g.RR + 2, g.RC + 2), GARBAGE)                # whatever lies beyond the region must not matter
            p[1:-1, 1:-1] = o[key]
            v[key] = 0.2 * ((((p[1:-1, 1:-1] + p[1:-1, :-2]) + p[1:-1, 2:]) + p[2:, 1:-1]) + p[:-2, 1:-1])
        if s % T == 0 and s < nsweeps:
            g.exchange({"v": v})
        for key, t in g.t.items():
            v[key] = np.where(t["interior"], v[key], fx[key][s & 1])
            if s >= nsweeps - 1:
                (outB if s & 1 else outA)[t["gi"][t["own"]], t["gj"][t["own"]]] = v[key][t["own"]]
            o[key] = v[key]
    return outA, outB


def emulate_fdtd(ex, ey, hz, fict, rb, cb, nw, T, PI, PJ):
    nx, ny = ex.shape
    tmax = len(fict)
    g = Tiles(nx, ny, 0, rb, cb, nw, T, PI, PJ)
    X, Y, Z = {}, {}, {}
    for key, t in g.t.items():
        gi, gj = t["gi"], t["gj"]
        inside = (gi >= 0) & (gi < nx) & (gj >= 0) & (gj < ny)
        t["inside"] = inside
        t["halo"] = inside & ~t["own"] & t["need"]
        ci, cj = np.clip(gi, 0, nx - 1), np.clip(gj, 0, ny - 1)
        X[key], Y[key], Z[key] = (np.where(inside, f[ci, cj], 0.0) for f in (ex, ey, hz))
    for s in range(1, tmax + 1):
        for key, t in g.t.items():
            gi, gj, inside = t["gi"], t["gj"], t["inside"]
            x, y, z = X[key], Y[key], Z[key]
            zp = np.full((g.RR + 2, g.RC + 2), GARBAGE); zp[1:-1, 1:-1] = z
            yn = np.where(inside, y - 0.5 * (z - zp[:-2, 1:-1]), y)
            yn = np.where(inside & (gi == 0), fict[s - 1], yn)
            xn = np.where(inside & (gj != 0), x - 0.5 * (z - zp[1:-1, :-2]), x)
            xp = np.full((g.RR + 2, g.RC + 2), GARBAGE); xp[1:-1, 1:-1] = xn
            yp = np.full((g.RR + 2, g.RC + 2), GARBAGE); yp[1:-1, 1:-1] = yn
            zn = z - 0.7 * (((xp[1:-1, 2:] - xn) + yp[2:, 1:-1]) - yn)
            zn = np.where(inside & (gi != nx - 1) & (gj != ny - 1), zn, z)
            X[key], Y[key], Z[key] = xn, yn, zn
        if s % T == 0 and s < tmax:
            g.exchange({"x": X, "y": Y, "z": Z})
    out = [f.copy() for f in (ex, ey, hz)]
    for key, t in g.t.items():
        for f, reg in zip(out, (X[key], Y[key], Z[key])):
            f[t["gi"][t["own"]], t["gj"][t["own"]]] = reg[t["own"]]
    return out


@pytest.mark.parametrize("ts,shape,sms", [(6, (150, 150), 148), (8, (97, 450), 12), (9, (257, 339), 148), (12, (40, 13), 2),
                                          (3, (3, 3), 148), (4, (300, 5), 9), (5, (113, 114), 6), (2, (200, 200), 148),
                                          (14, (6, 6), 1), (10, (31, 64), 3)])
def test_jacobi2d_scheme_reproduces_the_oracle(ts, shape, sms):
    p = plan_j2(ts, shape[0], shape[1], sms)
    if p is None:
        pytest.skip("grid does not run resident on %d SMs" % sms)
    rng = np.random.default_rng(ts * 7 + shape[1])
    A, B = rng.random(shape) - 0.5, rng.random(shape) - 0.5
    gA, gB = emulate_jacobi(A, B, ts, *p)
    oracle.jacobi_2d(ts, A, B)
    assert np.array_equal(gA, A) and np.array_equal(gB, B)


@pytest.mark.parametrize("cfg", [(2, 2, 3, 1), (4, 2, 4, 2), (4, 2, 8, 3), (8, 2, 3, 2), (4, 4, 6, 3)])
def test_jacobi2d_scheme_with_forced_configurations(cfg):
    rb, cb, nw, T = cfg
    shape, ts = (113, 170), 7
    cap_i, cap_j = nw * rb - 2 * T, 32 * cb - 2 * T
    PI, PJ = -(-(shape[0] - 2) // cap_i), -(-(shape[1] - 2) // cap_j)
    check_plan(shape[0] - 2, shape[1] - 2, PI * PJ, rb, cb, nw, T, PI, PJ)
    rng = np.random.default_rng(5)
    A, B = rng.random(shape) - 0.5, rng.random(shape) - 0.5
    gA, gB = emulate_jacobi(A, B, ts, rb, cb, nw, T, PI, PJ)
    oracle.jacobi_2d(ts, A, B)
    assert np.array_equal(gA, A) and np.array_equal(gB, B)


@pytest.mark.parametrize("tm,shape,sms", [(6, (150, 150), 148), (8, (97, 450), 12), (9, (57, 139), 148), (12, (40, 13), 2),
                                          (3, (3, 3), 148), (4, (300, 5), 9), (5, (113, 114), 6), (1, (200, 220), 148),
                                          (11, (2, 300), 148), (6, (150, 1), 148)])
def test_fdtd2d_scheme_reproduces_the_oracle(tm, shape, sms):
    p = plan_f2(tm, shape[0], shape[1], sms)
    if p is None:
        pytest.skip("grid does not run resident on %d SMs" % sms)
    rng = np.random.default_rng(tm * 7 + shape[1])
    ex, ey, hz = rng.random(shape) - 0.5, rng.random(shape) - 0.5, rng.random(shape) - 0.5
    fict = rng.random(tm)
    got = emulate_fdtd(ex, ey, hz, fict, *p)
    oracle.fdtd_2d(tm, ex, ey, hz, fict)
    for a, w in zip(got, (ex, ey, hz)):
        assert np.array_equal(a, w)
